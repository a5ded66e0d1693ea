#!/bin/bash
# session 2, call 9: main stream = rest kernels only, as programmatic dependent launches; device-side wait for trsm(d); marker stream for the events
O=gpurun_out/s2c9; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 120 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], g['termination'], max(g['backward_errors'] or [0]))"
}
EXTRA="--config 2 --nodes 3000 --loops 400 --chains 2" run small_r1 PGS_REST_PDL=1
EXTRA="--config 2 --chains 2" run c2_r1 PGS_REST_PDL=1
EXTRA="--config 2 --chains 2" run c2_r0 PGS_REST_PDL=0
EXTRA="--config 3 --chains 2" run c3_r1 PGS_REST_PDL=1
EXTRA="--config 3 --chains 2" run c3_r0 PGS_REST_PDL=0
EXTRA="--config 3 --chains 2" run c3_r1_s140 PGS_REST_PDL=1 PGS_REST_SMS=140
EXTRA="--config 3 --chains 1" run c3_c1_r2 PGS_REST_PDL=2
EXTRA="--config 3 --chains 1" run c3_c1_r0 PGS_REST_PDL=0
PGS_REST_PDL=1 timeout 120 python tools/timeline_lab.py --config 3 --chains 2 > $O/timeline_c3_c2_r1.txt 2>$O/timeline_c3_c2_r1.err; tail -1 $O/timeline_c3_c2_r1.txt
PGS_REST_PDL=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x > $O/suite_part_r2.txt 2>&1; tail -3 $O/suite_part_r2.txt
