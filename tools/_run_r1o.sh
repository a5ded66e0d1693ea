mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky8.json 2> gpurun_out/solve_c3_sky8.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c3_sky8.json'))['gpu0']; print('c3', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky8.err
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 10000 -c 1200 --csv --log-file gpurun_out/launches_sky8_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky8.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky8_c3.csv
