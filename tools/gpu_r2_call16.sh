#!/bin/bash
mkdir -p gpurun_out/r2c16
O=gpurun_out/r2c16
timeout 1500 python -m pytest tests -m gpu -q -x > $O/gpu_suite.txt 2>&1; tail -6 $O/gpu_suite.txt
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c16/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'))
for k in ('lm','lm_c2','lm_c3_tight','lm_sharded'):
    s=d.get(k,{}); print(k, {x:s.get(x) for x in ('error','lm_iters_per_s','ms_total','final_cost','n_chains','switches_off','outliers','inliers_on_frac','outliers_off_frac')}, (s.get('linear_backward_error') or {}).get('max'))
print('e2e_trigger', d.get('e2e_trigger'))
print('roofline_large', d.get('lm_sharded',{}).get('roofline_large'))
print('cpu', d.get('cpu_baseline'))
PY
