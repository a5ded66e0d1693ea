#!/bin/bash
# Round check on a B200 box (run through gpurun): GPU tests, bench (both arms), launch list + full ncu capture of the
# sweep, launch list of the skyline factorisation.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
# round 1 ended with an xdist worker crash that was traced to a buffer overrun in pgs_facade_compose: confirm the fix, full output kept
python -m pytest tests -m gpu -q -n 4 --dist loadfile > gpurun_out/gpu_suite_xdist.txt 2>&1; tail -3 gpurun_out/gpu_suite_xdist.txt
python tools/facade_alternative_check.py 2>&1 | tail -4
# the reference's own PoseGraphSLAM.cpp with libpgs.so serving its ceres::Solve calls, against the same served by the oracle
python tests/reference_node_with_libpgs.py 2>/dev/null | grep 'wake-up' 
python tools/fourdof_bench.py > gpurun_out/fourdof_bench.txt 2>&1; cat gpurun_out/fourdof_bench.txt
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 3 -f -o gpurun_out/sweep_full python bench.py --steps 2 --warmup 3 --no-lm --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 10000 -c 1200 --csv --log-file gpurun_out/launches_skyline_c3.csv python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky.log 2>&1
python tools/launch_summary.py gpurun_out/launches_skyline_c3.csv
