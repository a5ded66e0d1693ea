mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_solve.py --config 2 --out gpurun_out/dist2_c2_v3.json > gpurun_out/dist2_c2_v3.log 2>&1
tail -c 1500 gpurun_out/dist2_c2_v3.log
python -c "
import json; d=json.load(open('gpurun_out/dist2_c2_v3.json')); print(d['dist']['ms_total'], d['single']['ms_total'], d['dist_vs_single'])"
