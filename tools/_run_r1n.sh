mkdir -p gpurun_out
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:sky_update -s 6000 -c 4 -f -o gpurun_out/sky_upd_full_r1n python tools/solve_bench.py --config 3 --solver skyline --max-iters 1 > gpurun_out/ncu_sky_full.log 2>&1
tail -3 gpurun_out/ncu_sky_full.log
