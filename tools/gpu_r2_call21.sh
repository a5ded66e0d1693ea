#!/bin/bash
mkdir -p gpurun_out/r2c21
O=gpurun_out/r2c21
./tools/bin/diag_lab 0 | tee $O/diag_lab_mode0.txt
./tools/bin/diag_lab 1 | tee $O/diag_lab_mode1.txt
run() { name=$1; shift
  env "$@" python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--config 3 --chains 1" run c3_c1_d0 PGS_DIAG_MODE=0
EXTRA="--config 3 --chains 1" run c3_c1_d1 PGS_DIAG_MODE=1
EXTRA="--config 3 --chains 2" run c3_c2_d0 PGS_DIAG_MODE=0
EXTRA="--config 3 --chains 2" run c3_c2_d1 PGS_DIAG_MODE=1
EXTRA="--config 2 --chains 2" run c2_c2_d0 PGS_DIAG_MODE=0
EXTRA="--config 2 --chains 2" run c2_c2_d1 PGS_DIAG_MODE=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_edge_cases_gpu.py -m gpu -q > $O/suite_part.txt 2>&1; tail -4 $O/suite_part.txt
