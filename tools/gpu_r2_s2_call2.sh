#!/bin/bash
# session 2, call 2: rest(d) with an equal number of tiles per CTA and as a programmatic dependent launch of rest(d-1)
O=gpurun_out/s2c2; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--config 3 --chains 2" run c3_b0p0 PGS_REST_BALANCE=0 PGS_REST_PDL=0
EXTRA="--config 3 --chains 2" run c3_b1p0 PGS_REST_BALANCE=1 PGS_REST_PDL=0
EXTRA="--config 3 --chains 2" run c3_b0p1 PGS_REST_BALANCE=0 PGS_REST_PDL=1
EXTRA="--config 3 --chains 2" run c3_b1p1 PGS_REST_BALANCE=1 PGS_REST_PDL=1
EXTRA="--config 3 --chains 2" run c3_b1p1_s140 PGS_REST_SMS=140
EXTRA="--config 3 --chains 2" run c3_b1p1_s148 PGS_REST_SMS=148
EXTRA="--config 3 --chains 2" run c3_b1p1_s124 PGS_REST_SMS=124
EXTRA="--config 3 --chains 2" run c3_b1p0_s148 PGS_REST_SMS=148 PGS_REST_PDL=0
EXTRA="--config 3 --chains 1" run c3_c1_b1p1 PGS_REST_SMS=132
EXTRA="--config 3 --chains 1" run c3_c1_b1p1_s148 PGS_REST_SMS=148
EXTRA="--config 2 --chains 2" run c2_b1p1 PGS_REST_SMS=132
timeout 300 python tools/timeline_lab.py --config 3 > $O/timeline_c3.txt 2>$O/timeline_c3.err; tail -1 $O/timeline_c3.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py -m gpu -q -x > $O/suite_part.txt 2>&1; tail -3 $O/suite_part.txt
