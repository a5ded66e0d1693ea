#!/bin/bash
# Round-2 second GPU call: the new chain / rank / baseline-size tests, the c3 solve with one and two chains, the full bench.
mkdir -p gpurun_out/r2c2
O=gpurun_out/r2c2
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_suite.txt 2>&1; tail -25 $O/gpu_suite.txt
python tests/reference_node_with_libpgs.py > $O/reference_node.txt 2>&1; tail -4 $O/reference_node.txt
python tools/solve_bench.py --config 3 --chains 1 > $O/solve_c3_chains1.json 2>$O/solve_c3_chains1.err; cut -c1-900 $O/solve_c3_chains1.json
python tools/solve_bench.py --config 3 --chains 2 > $O/solve_c3_chains2.json 2>$O/solve_c3_chains2.err; cut -c1-900 $O/solve_c3_chains2.json
python tools/solve_bench.py --config 3 --chains 4 > $O/solve_c3_chains4.json 2>$O/solve_c3_chains4.err; cut -c1-900 $O/solve_c3_chains4.json
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; cut -c1-6000 $O/bench.json
