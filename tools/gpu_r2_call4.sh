#!/bin/bash
mkdir -p gpurun_out/r2c4
O=gpurun_out/r2c4
python tools/parity_debug.py > $O/parity_debug.txt 2>&1; tail -22 $O/parity_debug.txt
for mode in 0 1 2; do
  PGS_UPDATE_MODE=$mode python tools/solve_bench.py --config 3 --chains 1 --max-iters 3 > $O/solve_c3_mode$mode.json 2>$O/solve_c3_mode$mode.err; python -c "
import json;g=json.load(open('$O/solve_c3_mode$mode.json'))['gpu0'];print('mode',$mode,'chains 1', g['ms_total'], g['ms_linear_solve'], g['final_cost'], g['backward_errors'])"
done
for sms in 148 140 124; do
  PGS_REST_SMS=$sms PGS_UPDATE_MODE=1 python tools/solve_bench.py --config 3 --chains 1 --max-iters 3 > $O/solve_c3_sms$sms.json 2>/dev/null; python -c "
import json;g=json.load(open('$O/solve_c3_sms$sms.json'))['gpu0'];print('sms',$sms,'chains 1', g['ms_total'], g['ms_linear_solve'], g['final_cost'])"
done
PGS_UPDATE_MODE=1 python tools/solve_bench.py --config 3 --chains 2 --max-iters 3 > $O/solve_c3_mode1_ch2.json 2>/dev/null; python -c "
import json;g=json.load(open('$O/solve_c3_mode1_ch2.json'))['gpu0'];print('mode 1 chains 2', g['ms_total'], g['ms_linear_solve'], g['final_cost'])"
PGS_UPDATE_MODE=2 python tools/solve_bench.py --config 3 --chains 2 --max-iters 3 > $O/solve_c3_mode2_ch2.json 2>/dev/null; python -c "
import json;g=json.load(open('$O/solve_c3_mode2_ch2.json'))['gpu0'];print('mode 2 chains 2', g['ms_total'], g['ms_linear_solve'], g['final_cost'])"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py tests/test_edge_cases_gpu.py -m gpu -q > $O/suite_part.txt 2>&1; tail -8 $O/suite_part.txt
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 6000 -c 1200 --csv --log-file $O/launches_skyline_c3.csv python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_sky.log 2>&1
python tools/launch_summary.py $O/launches_skyline_c3.csv | tee $O/launches_skyline_c3.txt
