mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_ -s 9000 -c 8 -f -o gpurun_out/sky_full_r1e python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky_full.log 2>&1
tail -3 gpurun_out/ncu_sky_full.log
