#!/bin/bash
O=gpurun_out/s2c8; mkdir -p $O
timeout 300 python tools/trigger_lab.py --config 3 2> $O/trigger_lab_c3.txt; cat $O/trigger_lab_c3.txt
timeout 300 python tools/trigger_lab.py --config 2 2> $O/trigger_lab_c2.txt; tail -12 $O/trigger_lab_c2.txt
