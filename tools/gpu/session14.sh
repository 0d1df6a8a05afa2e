#!/bin/bash
# round-2 evidence of the build with the haplotype-pair kernels, compact read layouts and cross-batch pipelining.
# usage: session14.sh bench | ncu1 | ncu2      (three gpurun calls: a call may bring back at most 64 MiB)
set -u
O=gpurun_out/s14; mkdir -p $O
prof() { # name cfg regex skip count
  timeout 300 python tools/quick_bench.py --cfg $2 --iters 1 > $O/plain_$1.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -o $O/prof_$1 -f python tools/quick_bench.py --cfg $2 --iters 1 > $O/ncu_$1.log 2>&1
  echo "prof $1 rc=$?"
}
case "${1:-bench}" in
bench)
  python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest.log
  ( time python bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -n 4 $O/bench.err
  ( time python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
  B="python bench.py --steps 5 --warmup 3 --no-configs --no-dispatcher --no-cpu-baseline --preheat-s 0.05"
  $B > $O/bench_short.json 2> $O/bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $O/bench_launches.csv $B > $O/ncu_bench.log 2>&1; echo "launch list rc=$?"
  for c in c1 c3 c4; do python tools/quick_bench.py --cfg $c --iters 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 12 --csv --log-file $O/${c}_launches.csv python tools/quick_bench.py --cfg $c --iters 1 > /dev/null 2>&1; done
  python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s14/bench.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/s14/bench_ref.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ref',round(r['value'],1),'same_config',d['config']==r['config'],'parity',d['parity']['ok'])
for k,c in d['configs'].items(): print(k, round(c['value']), round(c['roofline']['frac'],3), round(c['e2e']['value']), c['parity']['ok'])
dd=d['e2e_dispatcher']; print('disp', round(dd['c3_stream']['value']), round(dd['c4']['value']), dd['ok'])
PY
  ;;
ncu1)
  prof c2 c2 phmm_f32a_tier2 3 1
  prof c4 c4 phmm_f32p_tier2 3 1
  ;;
ncu2)
  prof c3 c3 phmm_f32p_tier2 3 1
  prof c5 c5 'phmm_f64' 3 1
  ;;
esac
du -sh $O
