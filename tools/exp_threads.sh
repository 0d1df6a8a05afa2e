for t in 4 2 1; do
  taskset -c 0-3 python bench.py --gpus 2 --threads $t --no-cpu-baseline 2>gpurun_out/exp1_t$t.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('threads',$t,'value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e']['ms_min_median_max_rank0'])
"
done
nproc
