set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_solve.py --config 2 --nodes 1500 --loops 300 --oracle --out gpurun_out/dist2_small.json > gpurun_out/dist2_small.log 2>&1
tail -c 3000 gpurun_out/dist2_small.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_solve.py --config 2 --out gpurun_out/dist2_c2.json > gpurun_out/dist2_c2.log 2>&1
tail -c 3000 gpurun_out/dist2_c2.log
