mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_8gpu_v2.json 2> gpurun_out/bench_8gpu_v2.err
tail -2 gpurun_out/bench_8gpu_v2.err; cut -c1-1800 gpurun_out/bench_8gpu_v2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/dist_solve.py --config 5 --no-single --out gpurun_out/dist8_c5_v2.json > gpurun_out/dist8_c5_v2.log 2>&1
tail -c 300 gpurun_out/dist8_c5_v2.log
python -c "
import json; d=json.load(open('gpurun_out/dist8_c5_v2.json')); x=d['dist']; print('dist', x['ms_total'], x['ms_linear_solve'], x['final_cost'], x['n_lm'], x['termination']); print([ (r['n_border_nodes'], r['ms_linear_solve']) for r in d['ranks']][:3])"
