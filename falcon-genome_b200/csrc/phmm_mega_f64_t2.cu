// FP64 rerun kernel, general form, register tier 2.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f64_tier2, kTierF64T2, QUEUE, double, true, 0, 2, PHMM_F64_TIER2, PHMM_F64_TIER2_N, PHMM_CLASSDESC_F64)
}
