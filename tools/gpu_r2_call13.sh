#!/bin/bash
mkdir -p gpurun_out/r2c13
O=gpurun_out/r2c13
run() { name=$1; shift
  env "$@" python tools/solve_bench.py --config 3 --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--chains 1" run m2_c1_s124 PGS_UPDATE_MODE=2 PGS_REST_SMS=124
EXTRA="--chains 1" run m2_c1_s140 PGS_UPDATE_MODE=2 PGS_REST_SMS=140
EXTRA="--chains 2" run m2_c2_s132 PGS_UPDATE_MODE=2 PGS_REST_SMS=132
EXTRA="--chains 2" run m2_c2_s124 PGS_UPDATE_MODE=2 PGS_REST_SMS=124
EXTRA="--chains 2" run m2_c2_s116 PGS_UPDATE_MODE=2 PGS_REST_SMS=116
EXTRA="--chains 2" run m2_c2_s108 PGS_UPDATE_MODE=2 PGS_REST_SMS=108
EXTRA="--chains 2" run m1_c2_s124 PGS_UPDATE_MODE=1 PGS_REST_SMS=124
EXTRA="--chains 2" run m1_c2_s108 PGS_UPDATE_MODE=1 PGS_REST_SMS=108
EXTRA="--chains 3" run m2_c3_s120 PGS_UPDATE_MODE=2 PGS_REST_SMS=120
EXTRA="--chains 4" run m2_c4_s120 PGS_UPDATE_MODE=2 PGS_REST_SMS=120
