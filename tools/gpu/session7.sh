#!/bin/bash
set -u
O=gpurun_out/s7; mkdir -p $O
python tools/dispatch_probe.py --devices 1 --callers 1,2,4 2>&1 | tail -n 4
FCS_PHMM_TIMELINE=1 python tools/dispatch_probe.py --devices 1 --callers 1 --calls 3 > $O/tl1.log 2> $O/tl1.err; grep -o "end@[0-9.]*\|sized+chunked\[w-1 c0\]@[0-9.]*" $O/tl1.err | tail -n 12
FCS_PHMM_PACK_THREADS=8 python tools/dispatch_probe.py --devices 1 --callers 1,4 2>&1 | tail -n 2
python tools/nam_bench.py --clients 16 --regions-per-call 1,8,64 --seconds 3 > $O/nam16.log 2>&1; tail -n 5 $O/nam16.log
python tools/nam_bench.py --clients 32 --regions-per-call 1,8 --seconds 3 > $O/nam32.log 2>&1; tail -n 4 $O/nam32.log
