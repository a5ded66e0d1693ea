mkdir -p gpurun_out
./tools/bin/fp64_lab > gpurun_out/fp64_lab.txt 2>&1; head -8 gpurun_out/fp64_lab.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-lm > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -2 gpurun_out/bench_r1l.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1l.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
