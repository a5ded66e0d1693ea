"""Builds BASELINE.json's configurations as explicit problems (nodes, SixDOFError blocks, switchable loop
blocks, regulariser) by running the facade's first trigger on the host (dry run): the same rules the
reference applies at src/PoseGraphSLAM.cpp:1381-1850 — odometry edges (u, u-f) with weight
0.9^f * exp(-yaw_deg^2/6), initial guesses, one regulariser per set root.  Used by bench.py and the tests."""
import numpy as np

from . import facade, synth

# fan-out used for BASELINE.json's edge counts (30k / 300k / 3M odometry edges = 3 per node); the
# reference's own literal is 5 (PoseGraphSLAM.cpp:1577).
CONFIG_FANOUT = {1: 1, 2: 3, 3: 3, 4: 3, 5: 3}


def build_problem(config=3, fanout=None, **spec_overrides):
    fanout = CONFIG_FANOUT[config] if fanout is None else fanout
    g = synth.generate_config(config, **spec_overrides)
    F = facade.Facade(odom_fanout=fanout, dry_run=True)
    F.ingest(g)
    if not F.solve_once():
        raise RuntimeError("facade did not trigger")
    o = F.odom_terms(); r = F.reg_terms(); q, t = F.poses()
    # loop edges whose endpoints lie in a dead zone are skipped by the trigger (PoseGraphSLAM.cpp:1400)
    ww = np.array([F.which_world(s) for s in g["stamps"]]) if len(g["k0"]) else np.zeros(g["N"], int)
    keep = (ww[g["la"]] >= 0) & (ww[g["lb"]] >= 0) if len(g["la"]) else np.zeros(0, bool)
    p = dict(config=config, fanout=fanout, N=g["N"], q=q, t=t, gt_q=g["gt_q"], gt_t=g["gt_t"],
             oc1=o["u"], oc2=o["umf"], oq=o["q"], ot=o["t"], ow=o["w"],
             la=g["la"][keep], lb=g["lb"][keep], lq=g["lq"][keep], lt=g["lt"][keep], lw=g["lw"][keep], lout=g["lout"][keep],
             rn=r["node"], rq=r["q"], rt=r["t"], rw=r["w"])
    F.close()
    return p


def shard_problem(p, rank, world):
    """Node-range sharding: rank k owns nodes [kN/P, (k+1)N/P); an edge belongs to the shard of
    its lower-index endpoint.  The shard keeps only the poses its blocks touch — its own range plus the halo of
    remote endpoints (at most the loop-gap bound above the range) — re-indexed locally in ascending global
    order (`nodes` maps local -> global), so a rank uploads and reads ~N/P poses, not N.  The sweep needs no
    exchange.  Returns a problem with only this rank's residual blocks."""
    if world == 1:
        return p
    N = p["N"]
    lo, hi = rank * N // world, (rank + 1) * N // world
    s = dict(p)
    om = np.minimum(p["oc1"], p["oc2"]); ok = (om >= lo) & (om < hi)
    lm = np.minimum(p["la"], p["lb"]); lk = (lm >= lo) & (lm < hi)
    for k in ("oc1", "oc2", "oq", "ot", "ow"):
        s[k] = p[k][ok]
    for k in ("la", "lb", "lq", "lt", "lw", "lout"):
        s[k] = p[k][lk]
    rk = (p["rn"] >= lo) & (p["rn"] < hi)
    for k in ("rn", "rq", "rt", "rw"):
        s[k] = p[k][rk]
    nodes = np.unique(np.concatenate([s["oc1"], s["oc2"], s["la"], s["lb"], s["rn"], np.arange(lo, hi)])).astype(np.int64)
    g2l = np.full(N, -1, np.int64); g2l[nodes] = np.arange(len(nodes))
    for k in ("oc1", "oc2", "la", "lb", "rn"):
        s[k] = g2l[s[k]].astype(np.int32)
    for k in ("q", "t", "gt_q", "gt_t"):
        s[k] = np.ascontiguousarray(p[k][nodes])
    s["N"] = len(nodes); s["nodes"] = nodes; s["n_halo"] = int(len(nodes) - (hi - lo))
    s["shard"] = (rank, world, lo, hi)
    return s


def load_into_solver(p, **options):
    from .capi import PoseGraphSolver
    S = PoseGraphSolver(**options)
    S.set_nodes(p["q"], p["t"])
    if len(p["oc1"]):
        S.add_odom_edges(p["oc1"], p["oc2"], p["oq"], p["ot"], p["ow"])
    if len(p["la"]):
        S.add_loop_edges(p["la"], p["lb"], p["lq"], p["lt"], p["lw"])
    if len(p["rn"]):
        S.set_regularizers(p["rn"], p["rq"], p["rt"], p["rw"])
    return S
