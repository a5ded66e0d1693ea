#!/bin/bash
# session 2, call 11: time per panel of ONE chain against the length of the chain (config-5 recipe cut to size)
O=gpurun_out/s2c11; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 200 python tools/solve_bench.py --max-iters 2 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;d=json.load(open('$O/$name.json'));g=d['gpu0'];print('$name', 'N', d['N'], 'ms_linear', round(g['ms_linear_solve'],1), 'us/panel/solve', round(g['ms_linear_solve']*1e3/2/(d['N']/16),2), g['final_cost'])"
}
EXTRA="--config 5 --nodes 125000 --loops 62500 --chains 1" run n125k_m1 PGS_CHAIN_MODE=1
EXTRA="--config 5 --nodes 250000 --loops 125000 --chains 1" run n250k_m1 PGS_CHAIN_MODE=1
EXTRA="--config 5 --nodes 500000 --loops 250000 --chains 1" run n500k_m1 PGS_CHAIN_MODE=1
EXTRA="--config 5 --nodes 500000 --loops 250000 --chains 1" run n500k_m0 PGS_CHAIN_MODE=0
EXTRA="--config 5 --nodes 500000 --loops 250000 --chains 1" run n500k_m1_pdl0 PGS_CHAIN_MODE=1 PGS_BACKWARD_PDL=0
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
