#!/bin/bash
# tests with the full-size parity cases + ncu evidence of this build (full captures of the three dominant kernels, launch list of a short bench)
set -u
O=gpurun_out/s11; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 $O/pytest.log
prof() { # name cfg regex skip count
  timeout 300 python tools/quick_bench.py --cfg $2 --iters 1 > $O/plain_$1.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -o $O/prof_$1 python tools/quick_bench.py --cfg $2 --iters 1 > $O/ncu_$1.log 2>&1
  echo "prof $1 rc=$?"
}
prof c2 c2 phmm_f32a_tier2 3 1
prof c4 c4 phmm_f32u_tier1 3 1
prof c5 c5 'phmm_f64' 3 1
prof c3 c3 'phmm_f32u_tier' 9 3
B="python bench.py --steps 5 --warmup 3 --no-configs --no-dispatcher --no-cpu-baseline --preheat-s 0.05"
$B > $O/bench_short.json 2> $O/bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $O/bench_launches.csv $B > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
ls -la $O | head -30
