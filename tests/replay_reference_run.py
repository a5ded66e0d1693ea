"""Replay a session RECORDED BY THE REFERENCE NODE and compare with what the reference itself optimised (SURVEY §8f-3:
"first real-data parity").  Test infrastructure: it uses the oracle as its checker, so it lives under tests/.

The reference writes, at shutdown, `log_posegraph.json` (NodeDataManager::saveAsJSON, src/NodeDataManager.cpp:503-628:
odometry keyframes, loop edges, kidnap stamps) and `log_optimized_poses.json` (PoseGraphSLAM::saveAsJSON,
src/PoseGraphSLAM.cpp:1111-1207: `wTc_opt` per keyframe, `switching_var_after_opt` per loop edge).  Given the directory
that holds them this script

  1. loads the recorded graph (product loader `pgs_facade_load_posegraph_json`, or a plain-Python reader for --oracle),
  2. solves it — on the GPU through the facade (default) or with the CPU oracle front-end + LM (--oracle) — as ONE trigger
     (the reference solved it in many wake-ups whose timing is not recorded, so agreement is expected at the level of
     "same minimum of the final problem", hence --tight, not of a 10-iteration trajectory),
  3. evaluates the REFERENCE's recorded solution under the oracle's restatement of the cost (what the real Ceres run ended
     with, as seen by our functors) and reports translation / rotation deviations per keyframe and the switch states.

    python tests/replay_reference_run.py <dir> [--oracle] [--fanout 5] [--tight] [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend, pgo  # noqa: E402


def parse_mat(s):
    """'a,b,c,d;e,...' (PoseManipUtils.cpp:272-295) -> 4x4."""
    rows = [[float(x) for x in r.split(",")] for r in s.strip().strip(";").split(";")]
    M = np.array(rows)
    if M.shape != (4, 4):
        raise ValueError(f"not a 4x4 matrix string: {s[:60]!r}")
    return M


def read_posegraph(directory):
    J = json.load(open(os.path.join(directory, "log_posegraph.json")))
    nodes, edges = J["nodes"] or [], J["loopedges"] or []
    T = np.array([parse_mat(n["wTc"]) for n in nodes]).reshape(-1, 4, 4)
    stamps = np.array([int(round(float(n["timestamp"]) * 1e9)) for n in nodes], np.int64)
    qt = [pgo.mat4_to_pose(M) for M in T]
    k0, k1 = [], []
    for k in (J.get("kidnap_info") or []):          # the reference writes null when there was no kidnap
        k0.append(int(k["stampNSec_started"]) if "stampNSec_started" in k else int(round(float(k["stamp_of_kidnap_i_started"]) * 1e9)))
        k1.append(int(k["stampNSec_ended"]) if "stampNSec_ended" in k else int(round(float(k["stamp_of_kidnap_i_ended"]) * 1e9)))
    bTa = [pgo.mat4_to_pose(parse_mat(e["b_T_a"])) for e in edges]
    return dict(N=len(nodes), stamps=stamps, q=np.array([x[0] for x in qt]).reshape(-1, 4), t=np.array([x[1] for x in qt]).reshape(-1, 3),
                k0=np.array(k0, np.int64), k1=np.array(k1, np.int64),
                la=np.array([e["idx0"] for e in edges], np.int32), lb=np.array([e["idx1"] for e in edges], np.int32),
                lq=np.array([x[0] for x in bTa]).reshape(-1, 4), lt=np.array([x[1] for x in bTa]).reshape(-1, 3),
                lw=np.array([float(e["weight"]) for e in edges]))


def read_optimized(directory):
    J = json.load(open(os.path.join(directory, "log_optimized_poses.json")))
    T = np.array([parse_mat(n["wTc_opt"]) for n in J["PoseGraphSLAM_nodes"]]).reshape(-1, 4, 4)
    sw = {int(e["getEdge_i"]): float(e["switching_var_after_opt"]) for e in (J.get("PoseGraphSLAM_loopedgeinfo") or []) if "switching_var_after_opt" in e}
    return T, sw


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def replay(directory, use_oracle=False, fanout=5, tight=False):
    g = read_posegraph(directory)
    T_ref, sw_ref = read_optimized(directory)
    opts = dict(max_num_iterations=200, function_tolerance=1e-14, parameter_tolerance=1e-12, gradient_tolerance=1e-12) if tight else {}
    # the oracle front-end always runs: it gives the residual blocks under which the reference's recorded solution is costed
    M = frontend.Manager(); M.ingest(g)
    R = frontend.ReferenceFrontEnd(M, odom_fanout=fanout, options=pgo.default_options(**opts) if opts else None)
    R.trigger(solve=False)
    P = R.problem()
    cost_initial = P.evaluate(jac=False)["cost"]
    n = min(len(T_ref), g["N"])
    qt = [pgo.mat4_to_pose(T_ref[i]) for i in range(n)]
    q_ref = np.array(R.opt_q).reshape(-1, 4).copy(); t_ref = np.array(R.opt_t).reshape(-1, 3).copy()
    q_ref[:n] = np.array([x[0] for x in qt]).reshape(-1, 4); t_ref[:n] = np.array([x[1] for x in qt]).reshape(-1, 3)
    Pr = R.problem(); Pr.set_nodes(q_ref, t_ref)
    if R.loops:
        Pr.set_switches(np.array([sw_ref.get(l[0], 0.99) for l in R.loops]))
    cost_reference = Pr.evaluate(jac=False)["cost"]
    if use_oracle:
        s = R.solve()
        q, t = np.array(R.opt_q).reshape(-1, 4), np.array(R.opt_t).reshape(-1, 3)
        sw = {l[0]: R.opt_s[l[0]] for l in R.loops}
        summary = dict(final_cost=s["final_cost"], iterations=len(s["iterations"]) - 1, termination=s["termination"], solver="oracle (CPU)")
    else:
        from solve_keyframe_pose_graph_b200 import facade
        F = facade.Facade(odom_fanout=fanout, **opts)
        F.load_posegraph_json(directory)
        if not F.solve_once(True):
            raise RuntimeError("the facade did not trigger a solve")
        q, t = F.poses(); F.n_loop = len(g["la"]); s_all = F.switches()
        sw = {l[0]: s_all[l[0]] for l in R.loops}
        sm = F.summary(); F.close()
        summary = dict(final_cost=sm["final_cost"], iterations=len(sm["iterations"]) - 1, termination=sm["termination"], solver="libpgs (GPU)")
    dt = np.array([np.linalg.norm(t[i] - T_ref[i][:3, 3]) for i in range(n)])
    dr = np.array([rot_angle(pgo.pose_to_mat4(q[i], t[i])[:3, :3], T_ref[i][:3, :3]) for i in range(n)])
    both = [e for e in sw if e in sw_ref]
    agree = sum((sw[e] > 0.5) == (sw_ref[e] > 0.5) for e in both)
    return dict(directory=directory, nodes=g["N"], loop_edges=len(g["la"]), blocks=dict(odometry=len(R.odom), loop=len(R.loops), regularisers=len(R.regs)),
                cost_at_odometry=cost_initial, cost_of_reference_solution=cost_reference, **summary,
                translation_dev_m=dict(max=float(dt.max()) if n else 0.0, median=float(np.median(dt)) if n else 0.0),
                rotation_dev_rad=dict(max=float(dr.max()) if n else 0.0, median=float(np.median(dr)) if n else 0.0),
                switches=dict(compared=len(both), same_state=int(agree)))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("directory")
    ap.add_argument("--oracle", action="store_true", help="solve with the CPU oracle instead of the GPU facade")
    ap.add_argument("--fanout", type=int, default=5, help="odometry fan-out (reference: 5, PoseGraphSLAM.cpp:1577)")
    ap.add_argument("--tight", action="store_true", help="iterate to convergence instead of the reference's 10-iteration cap")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    out = replay(a.directory, a.oracle, a.fanout, a.tight)
    print(json.dumps(out, indent=1))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
