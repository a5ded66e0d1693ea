#!/bin/bash
mkdir -p gpurun_out/suite
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/suite/gpu_suite.txt 2>&1; tail -40 gpurun_out/suite/gpu_suite.txt
