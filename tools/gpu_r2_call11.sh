#!/bin/bash
mkdir -p gpurun_out/r2c11
O=gpurun_out/r2c11
PGS_REST_SMS=140 ./tools/bin/ws_lab | tee $O/ws_lab.txt
run() { name=$1; shift
  env "$@" python tools/solve_bench.py --config 3 --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--chains 1" run m0_c1 PGS_UPDATE_MODE=0
EXTRA="--chains 1" run m1_c1_s140 PGS_UPDATE_MODE=1 PGS_REST_SMS=140
EXTRA="--chains 1" run m2_c1_s140 PGS_UPDATE_MODE=2 PGS_REST_SMS=140
EXTRA="--chains 1" run m2_c1_s132 PGS_UPDATE_MODE=2 PGS_REST_SMS=132
EXTRA="--chains 2" run m1_c2_s140 PGS_UPDATE_MODE=1 PGS_REST_SMS=140
EXTRA="--chains 2" run m1_c2_s132 PGS_UPDATE_MODE=1 PGS_REST_SMS=132
EXTRA="--chains 2" run m2_c2_s140 PGS_UPDATE_MODE=2 PGS_REST_SMS=140
EXTRA="--chains 2" run m2_c2_s132 PGS_UPDATE_MODE=2 PGS_REST_SMS=132
EXTRA="--chains 2" run m2_c2_s124 PGS_UPDATE_MODE=2 PGS_REST_SMS=124
PGS_UPDATE_MODE=1 PGS_REST_SMS=140 timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 6000 -c 1200 --csv --log-file $O/launches_skyline_c3_m1.csv python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_sky.log 2>&1
python tools/launch_summary.py $O/launches_skyline_c3_m1.csv | tee $O/launches_skyline_c3_m1.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_edge_cases_gpu.py -m gpu -q > $O/suite_part.txt 2>&1; tail -5 $O/suite_part.txt
