#!/bin/bash
# session 2: ncu --set full captures of the factorisation kernels in their last state (config 3, one chain so that the launches are not interleaved)
O=gpurun_out/s2ncu; mkdir -p $O
cap() { name=$1; regex=$2; skip=$3; n=$4
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$regex -s $skip -c $n -f -o $O/$name python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/$name.log 2>&1; tail -1 $O/$name.log | cut -c1-200
  python tools/ncu_summary.py $O/$name.ncu-rep "ncu --set full --clock-control none --cache-control none -k regex:$regex -s $skip -c $n, tools/solve_bench.py --config 3 --chains 1 --max-iters 1" > $O/$name.json; }
cap upd_ws "sky_update_ws_kernel" 600 4
cap diag2 "sky_diag2_kernel" 300 2
cap trsm "sky_trsm_kernel" 300 2
cap backward "sky_backward_kernel" 3000 3
ls -la $O | head -20
