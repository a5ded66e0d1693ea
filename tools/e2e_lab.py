#!/usr/bin/env python
"""End-to-end step (pgs_evaluate_from_host, pinned host buffers) timed by wall clock for the three upload variants,
interleaved in ONE process so that box-to-box and minute-to-minute drift hits them alike:
PGS_E2E_MODE=0 chunked cudaMemcpyAsync + partial sweeps, 1 chunked reads of pinned memory by the pack kernel + partial
sweeps, 2 one shot (three copies, pack, one sweep).  python tools/e2e_lab.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solve_keyframe_pose_graph_b200 import problems
p = problems.build_problem(3)
S = problems.load_into_solver(p)
q = torch.from_numpy(np.ascontiguousarray(p["q"])).pin_memory(); t = torch.from_numpy(np.ascontiguousarray(p["t"])).pin_memory()
s = torch.full((len(p["la"]),), 0.99, dtype=torch.float64).pin_memory()
ref = S.evaluate()["cost"]
res = {0: [], 1: [], 2: []}
for rep in range(12):
    for mode in (0, 1, 2):
        os.environ["PGS_E2E_MODE"] = str(mode)
        for _ in range(5): c = S.evaluate_from_host_ptr(q.data_ptr(), t.data_ptr(), s.data_ptr())
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(100): c = S.evaluate_from_host_ptr(q.data_ptr(), t.data_ptr(), s.data_ptr())
        torch.cuda.synchronize(); res[mode].append((time.perf_counter() - t0) / 100 * 1e6)
        assert c == ref
for mode, name in ((0, "chunked copies"), (1, "chunked reads of pinned memory"), (2, "one shot")):
    v = np.array(res[mode]); print(f"mode {mode} ({name}): median {np.median(v):.1f} us  min {v.min():.1f}  max {v.max():.1f}  (12 x 100 steps)")
