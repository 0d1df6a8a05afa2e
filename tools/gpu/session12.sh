#!/bin/bash
# ncu evidence of this build: full captures of the three dominant kernels (< 64 MiB in total), launch list of a short bench,
# and the host-thread trade-off on a 4-core affinity mask (what a rank gets on an 8-GPU / 32-core box)
set -u
O=gpurun_out/s12; mkdir -p $O
prof() { # name cfg regex skip count
  timeout 300 python tools/quick_bench.py --cfg $2 --iters 1 > $O/plain_$1.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -o $O/prof_$1 python tools/quick_bench.py --cfg $2 --iters 1 > $O/ncu_$1.log 2>&1
  echo "prof $1 rc=$?"
}
prof c2 c2 phmm_f32a_tier2 3 1
prof c4 c4 phmm_f32u_tier1 3 1
prof c5 c5 'phmm_f64' 3 1
B="python bench.py --steps 5 --warmup 3 --no-configs --no-dispatcher --no-cpu-baseline --preheat-s 0.05"
$B > $O/bench_short.json 2> $O/bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $O/bench_launches.csv $B > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
for t in 1 2 3 4; do echo "== 4 cores, $t packing threads"; FCS_PHMM_PACK_THREADS=$t taskset -c 0-3 python tools/quick_bench.py --cfg c2 --iters 3 --e2e 2>&1 | tail -n 1; done
for t in 2 4; do echo "== all cores, $t packing threads"; FCS_PHMM_PACK_THREADS=$t python tools/quick_bench.py --cfg c2 --iters 3 --e2e 2>&1 | tail -n 1; done
du -sh $O
