#!/bin/bash
# session 2, call 10: update kernel skips the warp pieces above the diagonal
O=gpurun_out/s2c10; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 120 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], g['termination'], max(g['backward_errors'] or [0]))"
}
EXTRA="--config 3 --chains 2" run c3_c2 PGS_X=0
EXTRA="--config 3 --chains 2" run c3_c2_again PGS_X=0
EXTRA="--config 3 --chains 1" run c3_c1 PGS_X=0
EXTRA="--config 2 --chains 2" run c2_c2 PGS_X=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x > $O/suite_part.txt 2>&1; tail -3 $O/suite_part.txt
