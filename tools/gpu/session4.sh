#!/bin/bash
# round-2: tests on the new build, per-config numbers, e2e timelines and knob comparisons, first full bench line
set -u
O=gpurun_out/s4; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for c in c2 c1 c3 c4 c5; do
  timeout 300 python tools/quick_bench.py --cfg $c --iters 10 --e2e > $O/qb_$c.log 2>&1; echo "qb $c rc=$?"; tail -2 $O/qb_$c.log
done
for c in c2 c3; do
  FCS_PHMM_TIMELINE=1 timeout 300 python tools/quick_bench.py --cfg $c --iters 2 --e2e > $O/tl_$c.log 2> $O/tl_$c.err
  for knob in FCS_PHMM_F64_SERIAL=1 FCS_PHMM_NO_TWO_PLANE=1 FCS_PHMM_CHUNKS_PER_THREAD_X10=15 FCS_PHMM_CHUNKS_PER_THREAD_X10=30 FCS_PHMM_PACK_THREADS=6; do
    echo "== $c $knob"; env $knob timeout 300 python tools/quick_bench.py --cfg $c --iters 3 --e2e 2>&1 | tail -1
  done
done
( time python bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err; head -c 1500 $O/bench.json
