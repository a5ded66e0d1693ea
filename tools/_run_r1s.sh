mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1s.json 2> gpurun_out/bench_r1s.err; tail -2 gpurun_out/bench_r1s.err; cat gpurun_out/bench_r1s.json | cut -c1-3000
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1s.json 2>&1; cut -c1-600 gpurun_out/bench_ref_r1s.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1s.csv python bench.py --steps 5 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench_r1s.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 3 -f -o gpurun_out/sweep_full_r1s python bench.py --steps 2 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 10000 -c 1200 --csv --log-file gpurun_out/launches_sky_final_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky_final.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky_final_c3.csv
timeout 600 python tools/solve_bench.py --config 2 --solver skyline --oracle > gpurun_out/solve_c2_final.json 2> gpurun_out/solve_c2_final.err
python -c "
import json; D=json.load(open('gpurun_out/solve_c2_final.json')); d=D['gpu0']; print('c2', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'], D.get('parity'), 'oracle wall', D['oracle']['wall_s'])"
timeout 600 python tools/solve_bench.py --config 4 --solver skyline > gpurun_out/solve_c4_final.json 2> gpurun_out/solve_c4_final.err
python -c "
import json; D=json.load(open('gpurun_out/solve_c4_final.json')); d=D['gpu0']; print('c4', D['N'], D['n_odom'], D['n_loop'], d['ms_linear_solve'], d['initial_cost'], d['final_cost'], d['lm_iters_per_s'], d['termination'])"
tail -2 gpurun_out/solve_c4_final.err
