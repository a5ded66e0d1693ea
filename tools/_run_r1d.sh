mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-lm --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -2 gpurun_out/bench_r1d.err; cut -c1-1200 gpurun_out/bench_r1d.json
timeout 600 python tools/solve_bench.py --config 2 --solver skyline > gpurun_out/solve_c2_sky3.json 2> gpurun_out/solve_c2_sky3.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c2_sky3.json'))['gpu0']; print('c2', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky3.json 2> gpurun_out/solve_c3_sky3.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c3_sky3.json'))['gpu0']; print('c3', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 1500 --csv --log-file gpurun_out/launches_sky3_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky3.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky3_c3.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 27000 -c 600 --csv --log-file gpurun_out/launches_sky3_c3_back.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky3b.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky3_c3_back.csv
