#!/bin/bash
mkdir -p gpurun_out/r2c6
O=gpurun_out/r2c6
PGS_UPDATE_MODE=1 PGS_REST_SMS=140 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:sky_update_ws_kernel -s 600 -c 3 -f -o $O/upd_ws python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_ws.log 2>&1; tail -2 $O/ncu_ws.log
PGS_UPDATE_MODE=0 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"sky_update_kernel<1>|sky_update_kernelILi1" -s 600 -c 3 -f -o $O/upd_old python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_old.log 2>&1; tail -2 $O/ncu_old.log
PGS_UPDATE_MODE=0 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:sky_diag_kernel -s 300 -c 2 -f -o $O/diag python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_diag.log 2>&1; tail -2 $O/ncu_diag.log
ls -la $O
