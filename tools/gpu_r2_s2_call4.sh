#!/bin/bash
# session 2, call 4: chain mode 1 = C(d) and trsm(d) alternate on the chain stream as programmatic dependent launches
O=gpurun_out/s2c4; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--config 3 --chains 1" run c3_c1_m0 PGS_CHAIN_MODE=0
EXTRA="--config 3 --chains 1" run c3_c1_m1 PGS_CHAIN_MODE=1
EXTRA="--config 3 --chains 2" run c3_c2_m0 PGS_CHAIN_MODE=0
EXTRA="--config 3 --chains 2" run c3_c2_m1 PGS_CHAIN_MODE=1
EXTRA="--config 2 --chains 2" run c2_c2_m0 PGS_CHAIN_MODE=0
EXTRA="--config 2 --chains 2" run c2_c2_m1 PGS_CHAIN_MODE=1
EXTRA="--config 2 --chains 1" run c2_c1_m0 PGS_CHAIN_MODE=0
EXTRA="--config 2 --chains 1" run c2_c1_m1 PGS_CHAIN_MODE=1
PGS_CHAIN_MODE=1 timeout 300 python tools/timeline_lab.py --config 3 --chains 1 > $O/timeline_c3_c1_m1.txt 2>$O/timeline_c3_c1_m1.err; tail -1 $O/timeline_c3_c1_m1.txt
PGS_CHAIN_MODE=1 timeout 300 python tools/timeline_lab.py --config 3 --chains 2 > $O/timeline_c3_c2_m1.txt 2>$O/timeline_c3_c2_m1.err; tail -1 $O/timeline_c3_c2_m1.txt
PGS_CHAIN_MODE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x > $O/suite_part_m1.txt 2>&1; tail -3 $O/suite_part_m1.txt
