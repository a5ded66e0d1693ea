#!/bin/bash
# end-to-end step against the number of upload chunks (PGS_E2E_CHUNKS), one process per setting, mode 0 only
O=gpurun_out/s2c15; mkdir -p $O
for c in 0.5 0.6 0.7 0.8 0.6 0.7; do
PGS_E2E_CHUNKS=2 PGS_E2E_SPLIT=$c python - <<PY 2>&1 | tee -a $O/e2e_split.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from solve_keyframe_pose_graph_b200 import problems
p = problems.build_problem(3); S = problems.load_into_solver(p)
q = torch.from_numpy(np.ascontiguousarray(p["q"])).pin_memory(); t = torch.from_numpy(np.ascontiguousarray(p["t"])).pin_memory()
s = torch.full((len(p["la"]),), 0.99, dtype=torch.float64).pin_memory()
ref = S.evaluate()["cost"]; res = []
for rep in range(10):
    for _ in range(5): c = S.evaluate_from_host_ptr(q.data_ptr(), t.data_ptr(), s.data_ptr())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): c = S.evaluate_from_host_ptr(q.data_ptr(), t.data_ptr(), s.data_ptr())
    torch.cuda.synchronize(); res.append((time.perf_counter() - t0) / 100 * 1e6); assert c == ref
print("2 chunks, first", os.environ["PGS_E2E_SPLIT"], "median %.1f us  min %.1f  max %.1f" % (np.median(res), min(res), max(res)))
PY
done
