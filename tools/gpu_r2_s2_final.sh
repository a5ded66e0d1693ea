#!/bin/bash
mkdir -p gpurun_out/s2final
O=gpurun_out/s2final
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_suite.txt 2>&1; tail -4 $O/gpu_suite.txt
python __graft_entry__.py smoke 2>&1 | tail -1
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $SAN --tool memcheck --print-limit 20 python tools/solve_bench.py --config 2 --max-iters 2 > $O/sanitizer_memcheck_c2.txt 2>&1; tail -2 $O/sanitizer_memcheck_c2.txt
timeout 900 $SAN --tool racecheck --print-limit 20 python tools/solve_bench.py --config 2 --nodes 4000 --loops 800 --max-iters 1 --chains 2 > $O/sanitizer_racecheck_c2s.txt 2>&1; tail -2 $O/sanitizer_racecheck_c2s.txt
timeout 600 $SAN --tool synccheck --print-limit 20 python tools/solve_bench.py --config 2 --nodes 4000 --loops 800 --max-iters 1 --chains 2 > $O/sanitizer_synccheck_c2s.txt 2>&1; tail -2 $O/sanitizer_synccheck_c2s.txt
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 6000 -c 1200 --csv --log-file $O/launches_skyline_c3.csv python tools/solve_bench.py --config 3 --chains 1 --max-iters 1 > $O/ncu_sky.log 2>&1
python tools/launch_summary.py $O/launches_skyline_c3.csv | tee $O/launches_skyline_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-lm --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/launch_summary.py $O/launches_bench.csv | tee $O/launches_bench.txt
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2final/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
for k in ('lm','lm_c2','lm_c3_tight','lm_sharded'):
    s=d.get(k,{}); print(k, {x:s.get(x) for x in ('error','lm_iters_per_s','ms_total','final_cost','n_chains')}, (s.get('linear_backward_error') or {}).get('max'))
print('roofline_large', {k:v for k,v in d.get('lm_sharded',{}).get('roofline_large',{}).items() if k in ('achieved','frac','kernel_ms')})
r=json.loads(open('gpurun_out/s2final/bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r.get('lm'), r.get('e2e_trigger'))
PY
