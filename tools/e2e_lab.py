#!/usr/bin/env python
"""End-to-end step (pgs_evaluate_from_host, pinned host buffers) timed by wall clock for the upload variants:
PGS_E2E_MODE=0 chunked cudaMemcpyAsync, 1 pinned memory read by the pack kernel.  python tools/e2e_lab.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solve_keyframe_pose_graph_b200 import problems
p = problems.build_problem(3)
S = problems.load_into_solver(p)
q = torch.from_numpy(np.ascontiguousarray(p["q"])).pin_memory(); t = torch.from_numpy(np.ascontiguousarray(p["t"])).pin_memory()
s = torch.full((len(p["la"]),), 0.99, dtype=torch.float64).pin_memory()
ref = S.evaluate()["cost"]
for _ in range(10): c = S.evaluate_from_host_ptr(q.data_ptr(), t.data_ptr(), s.data_ptr())
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): c = S.evaluate_from_host_ptr(q.data_ptr(), t.data_ptr(), s.data_ptr())
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
print(f"PGS_E2E_MODE={os.environ.get('PGS_E2E_MODE','1')}: {dt*1e6:.1f} us per step, cost {c!r} (resident {ref!r}) equal={c==ref}")
# pageable buffers
qn, tn, sn = p["q"].copy(), p["t"].copy(), np.full(len(p["la"]), 0.99)
for _ in range(5): c2 = S.evaluate_from_host(qn, tn, sn)
t0 = time.perf_counter()
for _ in range(50): c2 = S.evaluate_from_host(qn, tn, sn)
print(f"  pageable: {(time.perf_counter()-t0)/50*1e6:.1f} us per step equal={c2==ref}")
