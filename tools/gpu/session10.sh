#!/bin/bash
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
for c in c3 c1; do for knob in NONE=1 FCS_PHMM_SORT_REGIONS=0 FCS_PHMM_NO_COARSE=1; do echo "== $c $knob"; env $knob python tools/quick_bench.py --cfg $c --iters 5 --e2e 2>&1 | tail -n 1; done; done
python tools/dispatch_probe.py --devices 1 --callers 1,4 --calls 24 2>&1 | tail -n 2
FCS_PHMM_SORT_REGIONS=0 python tools/dispatch_probe.py --devices 1 --callers 1,4 --calls 24 2>&1 | tail -n 2
