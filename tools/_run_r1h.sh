mkdir -p gpurun_out
./tools/bin/diag_lab > gpurun_out/diag_lab.txt 2>&1; cat gpurun_out/diag_lab.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/solve_bench.py --config 2 --solver skyline --oracle > gpurun_out/solve_c2_sky6.json 2> gpurun_out/solve_c2_sky6.err
python -c "
import json; D=json.load(open('gpurun_out/solve_c2_sky6.json')); d=D['gpu0']; print('c2', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'], D.get('parity'))"
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky6.json 2> gpurun_out/solve_c3_sky6.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c3_sky6.json'))['gpu0']; print('c3', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky6.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 10000 -c 1200 --csv --log-file gpurun_out/launches_sky6_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky6.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky6_c3.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 27000 -c 600 --csv --log-file gpurun_out/launches_sky6_c3_back.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky6b.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky6_c3_back.csv
