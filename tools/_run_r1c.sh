mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-lm > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -2 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
