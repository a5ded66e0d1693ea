set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/solve_bench.py --config 2 --solver skyline > gpurun_out/solve_c2_sky2.json 2> gpurun_out/solve_c2_sky2.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c2_sky2.json'))['gpu0']; print('c2', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky2.json 2> gpurun_out/solve_c3_sky2.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c3_sky2.json'))['gpu0']; print('c3', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 1200 --csv --log-file gpurun_out/launches_sky2_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky2_c3.csv
