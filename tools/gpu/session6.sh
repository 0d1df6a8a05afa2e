#!/bin/bash
# 2-GPU box: in-process dispatcher test, bench at N=2 (replicas + dispatcher), bench N=1 for the dispatcher baseline
set -u
O=gpurun_out/s6; mkdir -p $O
nvidia-smi -L; nproc
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu or concurrent" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"; tail -n 4 $O/bench_n1.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -n 4 $O/bench_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/s6/bench_n{n}.json') if l.startswith('{')][-1])
        dd=d['e2e_dispatcher']
        print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d['parity']['ok'], 'disp c3',round(dd['c3_stream']['value']),dd['c3_stream']['ms_per_call'],dd['c3_stream']['chunks'],'c4',round(dd['c4']['value']),dd['c4']['ok'],dd['c3_stream']['ok'], d['e2e']['per_rank_ms_per_call']['rows'])
    except Exception as e: print(n,'ERR',e)
PY
