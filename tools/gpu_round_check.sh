#!/bin/bash
# Round check on a B200 box (run through gpurun): GPU tests, bench (both arms), launch list + full ncu capture of the
# sweep, launch list of the skyline factorisation.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 3 -f -o gpurun_out/sweep_full python bench.py --steps 2 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 10000 -c 1200 --csv --log-file gpurun_out/launches_skyline_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky.log 2>&1
python tools/launch_summary.py gpurun_out/launches_skyline_c3.csv
