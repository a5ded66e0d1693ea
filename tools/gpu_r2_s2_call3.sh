#!/bin/bash
# session 2, call 3: where the two-chain factorisation loses time: split point between the chains, grid of next(d)
O=gpurun_out/s2c3; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
export PGS_REST_BALANCE=0 PGS_REST_PDL=0
EXTRA="--config 3 --chains 2" run c3_base PGS_SPLIT_BIAS=1.0
EXTRA="--config 3 --chains 2" run c3_bias1.06 PGS_SPLIT_BIAS=1.06
EXTRA="--config 3 --chains 2" run c3_bias1.12 PGS_SPLIT_BIAS=1.12
EXTRA="--config 3 --chains 2" run c3_bias0.94 PGS_SPLIT_BIAS=0.94
EXTRA="--config 3 --chains 2" run c3_next16 PGS_NEXT_CTAS=16
EXTRA="--config 3 --chains 2" run c3_next24 PGS_NEXT_CTAS=24
EXTRA="--config 3 --chains 2" run c3_next32 PGS_NEXT_CTAS=32
EXTRA="--config 3 --chains 2" run c3_next24_s124 PGS_NEXT_CTAS=24 PGS_REST_SMS=124
EXTRA="--config 3 --chains 2" run c3_next16_s140 PGS_NEXT_CTAS=16 PGS_REST_SMS=140
