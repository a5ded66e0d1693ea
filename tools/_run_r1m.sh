mkdir -p gpurun_out
./tools/bin/diag_lab > gpurun_out/diag_lab.txt 2>&1; cat gpurun_out/diag_lab.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky7.json 2> gpurun_out/solve_c3_sky7.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c3_sky7.json'))['gpu0']; print('c3', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky7.err
