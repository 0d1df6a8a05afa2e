#!/bin/bash
# 8-GPU box: in-process dispatcher strong scaling 1/2/4/8 on ONE box, replicas + dispatcher through bench.py at N=8 and N=4
set -u
O=gpurun_out/n8; mkdir -p $O
nproc; nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -n 2
for d in 1 2 4 8; do python tools/dispatch_probe.py --devices $d --callers 4,8 --calls 40 --batches 20 2>&1 | tail -n 2; done | tee $O/probe_c3.log
for d in 1 2 4 8; do python tools/dispatch_probe.py --devices $d --callers 4 --calls 20 --workload c4 --batches 10 2>&1 | tail -n 1; done | tee $O/probe_c4.log
FCS_PHMM_CHUNKS_PER_THREAD_X10=10 python tools/dispatch_probe.py --devices 8 --callers 8 --calls 40 --batches 20 2>&1 | tail -n 1
FCS_PHMM_PACK_THREADS=3 python tools/dispatch_probe.py --devices 8 --callers 8 --calls 40 --batches 20 2>&1 | tail -n 1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --dispatcher-callers 8 ) > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; tail -n 4 $O/bench_n8.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-dispatcher ) > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
for n in (8,4):
    try:
        d=json.loads([l for l in open(f'gpurun_out/n8/bench_n{n}.json') if l.startswith('{')][-1])
        print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d['parity']['ok'],d['parity']['all_ranks_ok'],'threads',d['run']['host_pack_threads_per_rank'],d['run']['host_threads'])
        print('   per rank',[[round(x,2) for x in r] for r in d['e2e']['per_rank_ms_per_call']['rows']])
        dd=d.get('e2e_dispatcher')
        if dd: print('   disp c3',round(dd['c3_stream']['value']),dd['c3_stream']['ms_per_call'],dd['c3_stream']['chunks'],'c4',round(dd['c4']['value']),dd['c4']['ok'],dd['c3_stream']['ok'])
    except Exception as e: print(n,'ERR',e)
PY
