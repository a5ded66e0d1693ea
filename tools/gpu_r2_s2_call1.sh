#!/bin/bash
# session 2, call 1: backward sweep as programmatic dependent launches (on/off), factorisation timeline, host topology
O=gpurun_out/s2c1; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python tools/solve_bench.py --max-iters 3 $EXTRA > $O/$name.json 2>$O/$name.err; python -c "
import json;g=json.load(open('$O/$name.json'))['gpu0'];print('$name', round(g['ms_total'],1), round(g['ms_linear_solve'],1), g['final_cost'], max(g['backward_errors']))"
}
EXTRA="--config 3 --chains 2" run c3_pdl0 PGS_BACKWARD_PDL=0
EXTRA="--config 3 --chains 2" run c3_pdl1 PGS_BACKWARD_PDL=1
EXTRA="--config 2 --chains 2" run c2_pdl0 PGS_BACKWARD_PDL=0
EXTRA="--config 2 --chains 2" run c2_pdl1 PGS_BACKWARD_PDL=1
EXTRA="--config 3 --chains 1" run c3_c1_pdl1 PGS_BACKWARD_PDL=1
timeout 300 python tools/timeline_lab.py --config 3 > $O/timeline_c3.txt 2>$O/timeline_c3.err; tail -3 $O/timeline_c3.txt
timeout 300 python tools/timeline_lab.py --config 3 --chains 1 > $O/timeline_c3_c1.txt 2>$O/timeline_c3_c1.err; tail -1 $O/timeline_c3_c1.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chains_gpu.py -m gpu -q -x > $O/suite_part.txt 2>&1; tail -3 $O/suite_part.txt
(lscpu | head -25; numactl -H; nvidia-smi topo -m; cat /sys/bus/pci/devices/*/numa_node | sort | uniq -c; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective) > $O/host_topology.txt 2>&1
