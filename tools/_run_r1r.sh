mkdir -p gpurun_out
./tools/bin/upd_lab 3200 450 2>&1 | grep -E "^N="
./tools/bin/upd_lab 5000 800 2>&1 | grep -E "^N="
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/solve_bench.py --config 2 --solver skyline --oracle > gpurun_out/solve_c2_sky9.json 2> gpurun_out/solve_c2_sky9.err
python -c "
import json; D=json.load(open('gpurun_out/solve_c2_sky9.json')); d=D['gpu0']; print('c2', d['ms_linear_solve'], d['final_cost'], d['lm_iters_per_s'], D.get('parity'))"
timeout 900 python tools/solve_bench.py --config 3 --solver skyline --repeat 2 > gpurun_out/solve_c3_sky9.json 2> gpurun_out/solve_c3_sky9.err
python -c "
import json; D=json.load(open('gpurun_out/solve_c3_sky9.json'))
for k in ('gpu0','gpu1'):
    d=D[k]; print('c3', k, d['ms_linear_solve'], d['ms_total'], d['final_cost'], d['lm_iters_per_s'])"
tail -3 gpurun_out/solve_c3_sky9.err
