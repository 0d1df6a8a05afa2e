#!/bin/bash
# round-2 baseline session: GPU tests, per-config kernel/e2e numbers, full ncu captures of the three kernels under work
set -u
O=gpurun_out/s1; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
for c in c2 c1 c3 c4 c5; do
  timeout 300 python tools/quick_bench.py --cfg $c --iters 10 --e2e > $O/qb_$c.log 2>&1; echo "qb $c rc=$?"
done
prof() { # name cfg regex skip count
  timeout 300 python tools/quick_bench.py --cfg $2 --iters 1 > $O/plain_$1.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -o $O/prof_$1 python tools/quick_bench.py --cfg $2 --iters 1 > $O/ncu_$1.log 2>&1
  echo "prof $1 rc=$?"
}
prof c2 c2 phmm_f32a_tier2 3 1
prof c4 c4 phmm_f32u_tier1 3 1
prof c5 c5 'phmm_f64' 3 2
ls -la $O
tail -3 $O/qb_*.log
