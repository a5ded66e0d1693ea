#!/bin/bash
mkdir -p gpurun_out/r2d8b
O=gpurun_out/r2d8b
for cc in 2.0 2.5; do
PGS_CARRY_COST=$cc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 8 --steps 5 --warmup 3 --no-single --no-cpu-baseline > $O/bench_8gpu_cc$cc.json 2> $O/bench_8gpu_cc$cc.err; tail -2 $O/bench_8gpu_cc$cc.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('$O/bench_8gpu_cc$cc.json').read().strip().splitlines()[-1])
s=d.get('lm_sharded',{})
print('carry $cc', {k:s.get(k) for k in ('error','ms_total','lm_iters_per_s','border_nodes','final_cost')})
print([ (r['rank'], round(r['ms_eliminate']), round(r['ms_wait_in_border_allreduce']), round(r['ms_border_system']), r['n_interior_nodes']) for r in s.get('ranks',[])])
PY
done
