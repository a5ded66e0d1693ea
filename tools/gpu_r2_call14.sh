#!/bin/bash
mkdir -p gpurun_out/r2c14
O=gpurun_out/r2c14
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.txt 2>&1; tail -12 $O/gpu_suite.txt
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $SAN --tool memcheck --print-limit 20 python tools/solve_bench.py --config 2 --max-iters 2 > $O/sanitizer_memcheck_c2.txt 2>&1; tail -3 $O/sanitizer_memcheck_c2.txt
timeout 900 $SAN --tool racecheck --print-limit 20 python tools/solve_bench.py --config 2 --nodes 5000 --loops 1000 --max-iters 2 --chains 2 > $O/sanitizer_racecheck_c2s.txt 2>&1; tail -3 $O/sanitizer_racecheck_c2s.txt
timeout 600 $SAN --tool synccheck --print-limit 20 python tools/solve_bench.py --config 2 --nodes 5000 --loops 1000 --max-iters 2 --chains 2 > $O/sanitizer_synccheck_c2s.txt 2>&1; tail -3 $O/sanitizer_synccheck_c2s.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 3 -f -o $O/sweep_full python bench.py --steps 2 --warmup 3 --no-lm --no-cpu-baseline > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ncu -i $O/sweep_full.ncu-rep --page raw --csv > $O/sweep_full.csv 2>/dev/null
