#!/bin/bash
mkdir -p gpurun_out/r2d8
O=gpurun_out/r2d8
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 --no-single > $O/bench_8gpu.json 2> $O/bench_8gpu.err; tail -3 $O/bench_8gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d8/bench_8gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
s=d.get('lm_sharded',{})
print({k:s.get(k) for k in ('error','workload','ms_total','lm_iters_per_s','border_nodes','border_buffer_bytes','n_collectives','final_cost','dist_vs_single','single_gpu')})
print([ (r['rank'], round(r['ms_total']), round(r['ms_linear_solve']), round(r['ms_comm']), round(r['ms_assemble']), r['factor_nnz'], r['n_interior_nodes'], r['n_local_border_nodes']) for r in s.get('ranks',[])])
print(s.get('linear_backward_error'))
PY
