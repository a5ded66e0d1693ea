#!/bin/bash
mkdir -p gpurun_out/r2d2
O=gpurun_out/r2d2
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x > $O/test_dist_gpu.txt 2>&1; tail -5 $O/test_dist_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --sharded-nodes 200000 > $O/bench_2gpu_200k.json 2> $O/bench_2gpu_200k.err; tail -3 $O/bench_2gpu_200k.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d2/bench_2gpu_200k.json').read().strip().splitlines()[-1])
s=d.get('lm_sharded',{})
print({k:s.get(k) for k in ('error','workload','ms_total','lm_iters_per_s','border_nodes','border_buffer_bytes','n_collectives','final_cost','dist_vs_single','single_gpu')})
print([ (r['rank'], round(r['ms_total']), round(r['ms_linear_solve']), round(r['ms_comm']), r['factor_nnz'], r['n_interior_nodes'], r['n_local_border_nodes']) for r in s.get('ranks',[])])
print(s.get('linear_backward_error'))
PY
