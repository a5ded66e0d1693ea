#!/bin/bash
# Round-2 first GPU call: carry-over checks of round 1 and compute-sanitizer evidence.
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
python -m pytest tests -m gpu -x -q > $O/gpu_suite_serial.txt 2>&1; tail -3 $O/gpu_suite_serial.txt
timeout 900 python -m pytest tests -m gpu -q -n 4 --dist loadfile > $O/gpu_suite_xdist.txt 2>&1; tail -3 $O/gpu_suite_xdist.txt
python tools/facade_alternative_check.py > $O/facade_alternative.txt 2>&1; tail -4 $O/facade_alternative.txt
python tests/reference_node_with_libpgs.py > $O/reference_node.txt 2>&1; grep 'wake-up' $O/reference_node.txt | tail -8
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --print-limit 20 python __graft_entry__.py smoke > $O/sanitizer_memcheck_smoke.txt 2>&1; tail -4 $O/sanitizer_memcheck_smoke.txt
timeout 900 $SAN --tool racecheck --print-limit 20 python __graft_entry__.py smoke > $O/sanitizer_racecheck_smoke.txt 2>&1; tail -4 $O/sanitizer_racecheck_smoke.txt
timeout 900 $SAN --tool memcheck --print-limit 20 python tools/solve_bench.py --config 2 --max-iters 2 > $O/sanitizer_memcheck_c2.txt 2>&1; tail -4 $O/sanitizer_memcheck_c2.txt
timeout 900 $SAN --tool racecheck --print-limit 20 python tools/solve_bench.py --config 2 --nodes 3000 --loops 600 --max-iters 2 > $O/sanitizer_racecheck_c2s.txt 2>&1; tail -4 $O/sanitizer_racecheck_c2s.txt
timeout 600 $SAN --tool synccheck --print-limit 20 python tools/solve_bench.py --config 2 --nodes 3000 --loops 600 --max-iters 2 > $O/sanitizer_synccheck_c2s.txt 2>&1; tail -3 $O/sanitizer_synccheck_c2s.txt
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > $O/nvsmi.txt; cat $O/nvsmi.txt
