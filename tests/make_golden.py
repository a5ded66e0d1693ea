"""Generates tests/golden/*.npz from the CPU oracle (run here, committed).  The reference has no golden
vectors of its own, so these pin the oracle against regressions and give the GPU tests a
fixture that does not need the oracle at all.  Usage: python tests/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import frontend, pgo  # noqa: E402
from solve_keyframe_pose_graph_b200 import synth  # noqa: E402

CASES = {
    "config1": dict(config=1, kw={}, fan=1),
    "config2_small": dict(config=2, kw=dict(n_nodes=1500, n_loop=300), fan=3),
    "config3_small": dict(config=3, kw=dict(n_nodes=2000, n_loop=1000), fan=3),
}


def run_case(c):
    g = synth.generate_config(c["config"], **c["kw"])
    M = frontend.Manager(); M.ingest(g)
    R = frontend.ReferenceFrontEnd(M, odom_fanout=c["fan"])
    R.trigger(solve=False)
    P = R.problem()
    e0 = P.evaluate(autodiff=True)
    s = P.solve()
    q, t = P.poses()
    it = s["iterations"]
    return dict(initial_cost=e0["cost"], final_cost=s["final_cost"], termination=s["termination"],
                iter_cost=np.array([r["cost"] for r in it]), iter_radius=np.array([r["trust_region_radius"] for r in it]),
                iter_success=np.array([r["step_is_successful"] for r in it]), iter_rho=np.array([r["relative_decrease"] for r in it]),
                q=q, t=t, switches=P.switches(), outlier=g["lout"],
                r_o_head=e0["r_o"][:16], J_o_head=e0["J_o"][:16], r_l_head=e0["r_l"][:16], J_l_head=e0["J_l"][:16],
                g_p_head=e0["g_p"][:16])


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, c in CASES.items():
        d = run_case(c)
        np.savez_compressed(os.path.join(out, name + ".npz"), **d)
        print(name, d["termination"], d["initial_cost"], d["final_cost"], len(d["iter_cost"]), "iters; switches off:", int((d["switches"] < 0.5).sum()),
              "outliers:", int(d["outlier"].sum()))
