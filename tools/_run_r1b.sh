mkdir -p gpurun_out
./tools/bin/sweep_lab 30 > gpurun_out/sweep_lab_a.txt 2>&1; cat gpurun_out/sweep_lab_a.txt
./tools/bin/fp64_lab > gpurun_out/fp64_lab.txt 2>&1; cat gpurun_out/fp64_lab.txt
