#!/bin/bash
set -u
O=gpurun_out/s13; mkdir -p $O
for t in 1 2 4; do echo "== 4 cores, $t packing threads"; FCS_PHMM_PACK_THREADS=$t taskset -c 0-3 python tools/quick_bench.py --cfg c2 --iters 3 --e2e 2>&1 | tail -n 1; done
for t in 2 4; do echo "== c3, 4 cores, $t packing threads"; FCS_PHMM_PACK_THREADS=$t taskset -c 0-3 python tools/quick_bench.py --cfg c3 --iters 3 --e2e 2>&1 | tail -n 1; done
( time python bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -n 4 $O/bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -n 4 $O/bench_ref.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s13/bench.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('gpurun_out/s13/bench_ref.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ref',round(r['value'],1),'same_config',d['config']==r['config'],'parity',d['parity']['ok'])
for k,c in d['configs'].items(): print(k, round(c['value']), round(c['roofline']['frac'],3), round(c['e2e']['value']), c['parity']['ok'])
dd=d['e2e_dispatcher']; print('disp', round(dd['c3_stream']['value']), round(dd['c4']['value']), dd['ok'])
PY
