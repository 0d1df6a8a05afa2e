#!/bin/bash
set -u
O=gpurun_out/s9; mkdir -p $O
nproc
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -n 2
python tools/dispatch_probe.py --devices 1 --callers 4 --calls 24 2>&1 | tail -n 1
python tools/dispatch_probe.py --devices 2 --callers 1,4,8 --calls 24 2>&1 | tail -n 3
python tools/dispatch_probe.py --devices 2 --callers 4 --calls 20 --workload c4 --batches 10 2>&1 | tail -n 1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -n 4 $O/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s9/bench_n2.json') if l.startswith('{')][-1])
dd=d['e2e_dispatcher']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d['parity']['ok'], 'disp c3',round(dd['c3_stream']['value']),dd['c3_stream']['ms_per_call'],dd['c3_stream']['chunks'],'c4',round(dd['c4']['value']),dd['c4']['ok'],dd['c3_stream']['ok'])
PY
