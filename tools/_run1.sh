set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/solve_bench.py --config 2 --solver skyline --oracle > gpurun_out/solve_c2_sky.json 2> gpurun_out/solve_c2_sky.err
tail -c 1500 gpurun_out/solve_c2_sky.json
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky.json 2> gpurun_out/solve_c3_sky.err
tail -c 1500 gpurun_out/solve_c3_sky.json; tail -3 gpurun_out/solve_c3_sky.err
