#!/bin/bash
# session 2: bench.py over N GPUs of one box (the sharded LM solve of config 5 rides on the line as lm_sharded).  usage: gpu_r2_s2_distN.sh N [extra bench flags]
N=$1; shift
O=gpurun_out/s2d$N; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 30 --warmup 5 "$@" > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; tail -2 $O/bench_${N}gpu.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('$O/bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
s=d.get('lm_sharded',{})
print({k:s.get(k) for k in ('error','ms_total','lm_iters_per_s','border_nodes','border_buffer_bytes','final_cost','dist_vs_single','single_gpu')})
print([ (r['rank'], round(r['ms_eliminate']), round(r['ms_linear_solve']), round(r['ms_wait_in_border_allreduce']), round(r['ms_border_system']), r['n_interior_nodes'], r['factor_nnz']) for r in s.get('ranks',[])])
print(s.get('linear_backward_error',{}).get('max'))
PY
