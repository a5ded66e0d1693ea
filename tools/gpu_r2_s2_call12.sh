#!/bin/bash
# session 2, call 12: hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) and the overlap of the streams of one solve, one process
O=gpurun_out/s2c12; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 120 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], g['termination'], max(g['backward_errors'] or [0]))"
}
EXTRA="--config 3 --chains 2" run c3_c2_conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
EXTRA="--config 3 --chains 2" run c3_c2_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
EXTRA="--config 3 --chains 2" run c3_c2_conn8b CUDA_DEVICE_MAX_CONNECTIONS=8
EXTRA="--config 3 --chains 2" run c3_c2_conn32b CUDA_DEVICE_MAX_CONNECTIONS=32
EXTRA="--config 3 --chains 1" run c3_c1_conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
EXTRA="--config 3 --chains 1" run c3_c1_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
EXTRA="--config 2 --chains 2" run c2_c2_conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
EXTRA="--config 2 --chains 2" run c2_c2_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 python tools/timeline_lab.py --config 3 --chains 2 > $O/timeline_c3_c2_conn32.txt 2>$O/timeline.err; tail -1 $O/timeline_c3_c2_conn32.txt; grep "mean durations" $O/timeline_c3_c2_conn32.txt
