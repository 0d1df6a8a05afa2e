"""falcon-genome_b200 — B200-native PairHMM forward-likelihood path for fcs-genome's
HaplotypeCaller / Mutect2 stages (C ABI library + thin host mirror).  See DESIGN.md."""
from .batch import FlatBatch, Region, partition_regions  # noqa: F401
from .pairhmm import PairHMM, PairHMMError, RegionArray, ResidentBatch, kernel_class, plan_check  # noqa: F401
from .capture import load_capture, read_gkl_text, save_capture, write_gkl_text  # noqa: F401,E402
from .prepost import finalize_region, prepare_read  # noqa: F401,E402
