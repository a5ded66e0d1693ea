#!/bin/bash
mkdir -p gpurun_out/r2c15
O=gpurun_out/r2c15
./tools/bin/ws_lab | tee $O/ws_lab.txt
run() { name=$1; shift
  env "$@" python tools/solve_bench.py --config 3 --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--chains 1" run c1
EXTRA="--chains 2" run c2
timeout 600 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q -k "100k or far_reaching" > $O/edge.txt 2>&1; tail -4 $O/edge.txt
