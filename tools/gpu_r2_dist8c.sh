#!/bin/bash
mkdir -p gpurun_out/r2d8c
O=gpurun_out/r2d8c
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29585 bench.py --gpus 8 --steps 30 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; tail -2 $O/bench_8gpu.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('$O/bench_8gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
s=d.get('lm_sharded',{})
print({k:s.get(k) for k in ('error','ms_total','lm_iters_per_s','border_nodes','border_buffer_bytes','final_cost','dist_vs_single','single_gpu')})
print([ (r['rank'], round(r['ms_eliminate']), round(r['ms_wait_in_border_allreduce']), round(r['ms_border_system']), r['n_interior_nodes'], r['factor_nnz']) for r in s.get('ranks',[])])
print(s.get('linear_backward_error'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29586 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/bench_8gpu_ref.json 2> $O/bench_8gpu_ref.err; cut -c1-200 $O/bench_8gpu_ref.json
