#!/bin/bash
mkdir -p gpurun_out/r2c7
O=gpurun_out/r2c7
run() { name=$1; shift
  env "$@" python tools/solve_bench.py --config 3 --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--chains 1" run m0_c1 PGS_UPDATE_MODE=0
EXTRA="--chains 2" run m0_c2 PGS_UPDATE_MODE=0
EXTRA="--chains 1" run m1_c1_s140 PGS_UPDATE_MODE=1 PGS_REST_SMS=140
EXTRA="--chains 1" run m2_c1_s140 PGS_UPDATE_MODE=2 PGS_REST_SMS=140
EXTRA="--chains 2" run m1_c2_s140 PGS_UPDATE_MODE=1 PGS_REST_SMS=140
EXTRA="--chains 2" run m1_c2_s132 PGS_UPDATE_MODE=1 PGS_REST_SMS=132
EXTRA="--chains 2" run m2_c2_s140 PGS_UPDATE_MODE=2 PGS_REST_SMS=140
EXTRA="--chains 2" run m2_c2_s132 PGS_UPDATE_MODE=2 PGS_REST_SMS=132
for mode in 0 1; do
PGS_UPDATE_MODE=$mode PGS_REST_SMS=140 timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 6000 -c 1200 --csv --log-file $O/launches_skyline_c3_m$mode.csv python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_sky.log 2>&1
python tools/launch_summary.py $O/launches_skyline_c3_m$mode.csv | tee $O/launches_skyline_c3_m$mode.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -x > $O/suite_part.txt 2>&1; tail -5 $O/suite_part.txt
