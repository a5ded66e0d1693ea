mkdir -p gpurun_out
./tools/bin/diag_lab > gpurun_out/diag_lab.txt 2>&1; cat gpurun_out/diag_lab.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-lm --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -2 gpurun_out/bench_r1f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['same_size_stream_write'])"
timeout 900 python tools/solve_bench.py --config 3 --solver skyline > gpurun_out/solve_c3_sky4.json 2> gpurun_out/solve_c3_sky4.err
python -c "
import json; d=json.load(open('gpurun_out/solve_c3_sky4.json'))['gpu0']; print('c3', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 1500 --csv --log-file gpurun_out/launches_sky4_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sky4_c3.csv
