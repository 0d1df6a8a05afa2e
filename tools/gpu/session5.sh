#!/bin/bash
set -u
O=gpurun_out/s5; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
for c in c2 c3; do
  for knob in NONE=1 FCS_PHMM_HS_COLS=320 FCS_PHMM_HS_COLS=400 FCS_PHMM_CHUNKS_PER_THREAD_X10=12 FCS_PHMM_PACK_THREADS=6 FCS_PHMM_PACK_THREADS=8; do
    echo "== $c $knob"; env $knob timeout 300 python tools/quick_bench.py --cfg $c --iters 5 --e2e 2>&1 | tail -n 2
  done
done
