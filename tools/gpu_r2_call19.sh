#!/bin/bash
mkdir -p gpurun_out/r2c19
O=gpurun_out/r2c19
run() { name=$1; shift
  env "$@" python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--config 3 --chains 1" run c3_c1_nograph PGS_BACKWARD_GRAPH=0
EXTRA="--config 3 --chains 1" run c3_c1 PGS_BACKWARD_GRAPH=1
EXTRA="--config 3 --chains 2" run c3_c2_nograph PGS_BACKWARD_GRAPH=0
EXTRA="--config 3 --chains 2" run c3_c2 PGS_BACKWARD_GRAPH=1
EXTRA="--config 2 --chains 2" run c2_c2_nograph PGS_BACKWARD_GRAPH=0
EXTRA="--config 2 --chains 2" run c2_c2 PGS_BACKWARD_GRAPH=1
EXTRA="--config 5 --chains 2" run c5_c2 PGS_BACKWARD_GRAPH=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py -m gpu -q > $O/suite_part.txt 2>&1; tail -3 $O/suite_part.txt
