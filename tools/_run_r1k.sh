mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -3 gpurun_out/bench_8gpu.err; cut -c1-1500 gpurun_out/bench_8gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/dist_solve.py --config 5 --out gpurun_out/dist8_c5.json > gpurun_out/dist8_c5.log 2>&1
tail -c 600 gpurun_out/dist8_c5.log
python -c "
import json; d=json.load(open('gpurun_out/dist8_c5.json')); x=d['dist']; print('dist', x['ms_total'], x['ms_linear_solve'], x['final_cost'], x['n_lm'], x['termination']); print([ (r['n_border_nodes'], r['factor_nnz'], r['ms_linear_solve']) for r in d['ranks']]); s=d.get('single'); print('single', s and (s['ms_total'], s['final_cost'])); print(d.get('dist_vs_single'))"
