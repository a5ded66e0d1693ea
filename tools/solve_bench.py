#!/usr/bin/env python
"""Solve one BASELINE config on the GPU and print the LM summary (per-phase device times) as JSON.
  python tools/solve_bench.py --config 3 --solver skyline [--oracle] [--max-iters 10]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--solver", default="skyline", choices=["skyline", "pcg"])
    ap.add_argument("--oracle", action="store_true", help="also run the CPU oracle and compare")
    ap.add_argument("--max-iters", type=int, default=10)
    ap.add_argument("--nodes", type=int, default=0)
    ap.add_argument("--loops", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--chains", type=int, default=0, help="elimination chains on the one GPU (0 = automatic)")
    args = ap.parse_args()
    import solve_keyframe_pose_graph_b200 as pgs
    from solve_keyframe_pose_graph_b200 import problems
    over = {}
    if args.nodes:
        over["n_nodes"] = args.nodes
    if args.loops:
        over["n_loop"] = args.loops
    t0 = time.perf_counter()
    p = problems.build_problem(args.config, **over)
    t_build = time.perf_counter() - t0
    out = {"config": args.config, "N": int(p["N"]), "n_odom": len(p["oc1"]), "n_loop": len(p["la"]), "outliers": int(p["lout"].sum()), "build_s": t_build}
    for rep in range(args.repeat):
        S = problems.load_into_solver(p, linear_solver=pgs.capi.SKYLINE_CHOLESKY if args.solver == "skyline" else pgs.capi.BLOCK_PCG,
                                      max_num_iterations=args.max_iters, chains=args.chains)
        t0 = time.perf_counter()
        s = S.solve()
        wall = time.perf_counter() - t0
        its = s.pop("iterations")
        s["wall_s"] = wall
        s["n_lm"] = len(its) - 1
        s["lm_iters_per_s"] = (len(its) - 1) / max(wall, 1e-9)
        s["costs"] = [r["cost"] for r in its]
        s["switches_off"] = int((S.switches() < 0.5).sum())
        s["backward_errors"] = [float(x) for x in S.linear_backward_errors()]
        out[f"gpu{rep}"] = s
        qs, ts = S.poses(); sw = S.switches()
        S.close()
    if args.oracle:
        from oracle import pgo
        pgo.build()
        from bench import oracle_problem
        P = oracle_problem(p)
        t0 = time.perf_counter()
        so = P.solve(pgo.default_options(max_num_iterations=args.max_iters))
        out["oracle"] = {"wall_s": time.perf_counter() - t0, "final_cost": so["final_cost"], "termination": so["termination"],
                         "costs": [r["cost"] for r in so["iterations"]]}
        qo, to = P.poses()
        out["parity"] = {"max_dt": float(np.abs(ts - to).max()),
                         "max_drot": float((2 * np.arccos(np.abs(np.sum(qs * qo, axis=1)).clip(0, 1))).max()),
                         "switch_states_equal": bool(np.array_equal(sw > 0.5, P.switches() > 0.5)),
                         "rel_cost": abs(s["final_cost"] - so["final_cost"]) / so["final_cost"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
