#!/usr/bin/env python
"""Multi-GPU LM solve, one process per GPU (launch with torchrun).  Every rank loads the same graph, attaches the
NCCL communicator through the C-ABI (pgs_dist_init) and calls pgs_solve; rank 0 additionally solves the same graph
on one GPU (and, for small graphs, with the CPU oracle) and reports the differences.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/dist_solve.py --config 2 [--nodes N --loops L] [--oracle] [--no-single]"""
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
    ap.add_argument("--nodes", type=int, default=0)
    ap.add_argument("--loops", type=int, default=0)
    ap.add_argument("--max-iters", type=int, default=10)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")   # plumbing only: broadcasts the NCCL id; the data path is NCCL inside libpgs
    import solve_keyframe_pose_graph_b200 as pgs
    from solve_keyframe_pose_graph_b200 import problems
    over = {}
    if args.nodes:
        over["n_nodes"] = args.nodes
    if args.loops:
        over["n_loop"] = args.loops
    p = problems.build_problem(args.config, **over)
    ids = [pgs.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    S = problems.load_into_solver(p, device=local_rank, max_num_iterations=args.max_iters)
    S.dist_init(rank, world, ids[0])
    dist.barrier()
    t0 = time.perf_counter()
    s = S.solve()
    wall = time.perf_counter() - t0
    qd, td = S.poses(); swd = S.switches()
    st = S.dist_stats()
    allst = [None] * world
    dist.all_gather_object(allst, dict(st, ms_total=s["ms_total"], ms_linear_solve=s["ms_linear_solve"], ms_sweep=s["ms_sweep"], ms_assemble=s["ms_assemble"], factor_nnz=s["factor_nnz"]))
    out = None
    if rank == 0:
        its = s.pop("iterations")
        out = {"world": world, "config": args.config, "N": int(p["N"]), "n_odom": len(p["oc1"]), "n_loop": len(p["la"]),
               "dist": dict(s, wall_s=wall, n_lm=len(its) - 1, costs=[r["cost"] for r in its], radius=[r["trust_region_radius"] for r in its],
                            switches_off=int((swd < 0.5).sum())),
               "ranks": allst}
    S.close()
    if rank == 0 and not args.no_single:
        T = problems.load_into_solver(p, device=local_rank, max_num_iterations=args.max_iters)
        t0 = time.perf_counter(); s1 = T.solve(); w1 = time.perf_counter() - t0
        q1, t1 = T.poses(); sw1 = T.switches(); T.close()
        it1 = s1.pop("iterations")
        out["single"] = dict(s1, wall_s=w1, n_lm=len(it1) - 1, costs=[r["cost"] for r in it1])
        out["dist_vs_single"] = {"max_dt": float(np.abs(td - t1).max()),
                                 "max_drot": float((2 * np.arccos(np.abs(np.sum(qd * q1, axis=1)).clip(0, 1))).max()),
                                 "max_dswitch": float(np.abs(swd - sw1).max()) if len(sw1) else 0.0,
                                 "switch_states_equal": bool(np.array_equal(swd > 0.5, sw1 > 0.5)),
                                 "rel_cost": abs(s["final_cost"] - s1["final_cost"]) / max(s1["final_cost"], 1e-300),
                                 "same_trajectory": [r["step_is_successful"] for r in its] == [r["step_is_successful"] for r in it1]}
    if rank == 0 and args.oracle:
        from oracle import pgo
        pgo.build()
        from bench import oracle_problem
        P = oracle_problem(p)
        so = P.solve(pgo.default_options(max_num_iterations=args.max_iters))
        qo, to = P.poses()
        out["dist_vs_oracle"] = {"max_dt": float(np.abs(td - to).max()),
                                 "max_drot": float((2 * np.arccos(np.abs(np.sum(qd * qo, axis=1)).clip(0, 1))).max()),
                                 "switch_states_equal": bool(np.array_equal(swd > 0.5, P.switches() > 0.5)),
                                 "rel_cost": abs(s["final_cost"] - so["final_cost"]) / max(so["final_cost"], 1e-300),
                                 "oracle_costs": [r["cost"] for r in so["iterations"]]}
    if rank == 0:
        print(json.dumps(out), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(out, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
