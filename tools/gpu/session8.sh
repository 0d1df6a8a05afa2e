#!/bin/bash
set -u
O=gpurun_out/s8; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
FCS_PHMM_DEBUG=1 python tools/dispatch_probe.py --devices 1 --callers 1,2,4 --calls 24 > $O/probe.log 2> $O/probe.err; cat $O/probe.log; grep -c "slot grows" $O/probe.err; grep "slot grows" $O/probe.err | tail -n 5
for c in c3 c1 c4; do echo "== $c"; python tools/quick_bench.py --cfg $c --iters 5 --e2e 2>&1 | tail -n 1; FCS_PHMM_RAMP=1 python tools/quick_bench.py --cfg $c --iters 5 --e2e 2>&1 | tail -n 1; done
