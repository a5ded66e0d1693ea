set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -2 gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1a.json 2>&1; cat gpurun_out/bench_ref_r1a.json
./tools/bin/sweep_lab 30 > gpurun_out/sweep_lab_a.txt 2>&1; cat gpurun_out/sweep_lab_a.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1a.csv python bench.py --steps 5 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench_r1a.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 3 -f -o gpurun_out/sweep_full_r1a python bench.py --steps 2 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
