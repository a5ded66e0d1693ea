"""Sharded linear solve on ONE GPU: several elimination chains inside one handle, and several
ranks of the multi-GPU scheme over the in-process transport (pgs_dist_init_local), each rank a handle on its own
host thread.  Every variant must walk the same LM trajectory as the plain natural-order skyline and end at the
oracle's poses (1e-5 m / 1e-4 rad, same switch states, cost 1e-5 relative — BASELINE.json north_star); between
two exact factorisations of the same system the bar is far tighter (1e-8 m)."""
import threading
import uuid

import numpy as np
import pytest

from util_graphs import load_oracle, load_pgs, random_graph, rot_angle_between

pytestmark = pytest.mark.gpu

import solve_keyframe_pose_graph_b200 as pgs  # noqa: E402
from solve_keyframe_pose_graph_b200 import problems  # noqa: E402


def _as_graph(p):
    return dict(N=p["N"], q=p["q"], t=p["t"], oc1=p["oc1"], oc2=p["oc2"], oq=p["oq"], ot=p["ot"], ow=p["ow"],
                la=p["la"], lb=p["lb"], lq=p["lq"], lt=p["lt"], lw=p["lw"], rn=p["rn"], rq=p["rq"], rt=p["rt"], rw=p["rw"])


def _solve(g, **opt):
    S = load_pgs(g, **opt)
    s = S.solve(); q, t = S.poses(); sw = S.switches()
    try:
        st = S.dist_stats()
    except pgs.PgsError:
        st = None
    be = S.linear_backward_errors()
    S.close()
    return s, q, t, sw, st, be


def _same_run(a, b, tol_t=1e-8, tol_r=1e-7):
    sa, qa, ta, swa = a[:4]; sb, qb, tb, swb = b[:4]
    assert [r["step_is_successful"] for r in sa["iterations"]] == [r["step_is_successful"] for r in sb["iterations"]]
    ca = np.array([r["cost"] for r in sa["iterations"]]); cb = np.array([r["cost"] for r in sb["iterations"]])
    assert np.abs(ca - cb).max() <= 1e-9 * np.abs(cb).max()
    assert sa["termination"] == sb["termination"]
    assert np.abs(ta - tb).max() < tol_t and rot_angle_between(qa, qb).max() < tol_r
    assert np.array_equal(swa > 0.5, swb > 0.5) and (np.abs(swa - swb).max() < 1e-7 if len(swa) else True)


@pytest.mark.parametrize("n,nl,seed", [(700, 120, 1), (2500, 600, 2)])
def test_chains_on_one_gpu_match_plain_and_oracle(n, nl, seed):
    g = random_graph(n, 3, nl, outlier_frac=0.1, seed=seed)
    plain = _solve(g, chains=1)
    assert plain[0]["n_chains"] == 1 and plain[4] is None
    for chains in (2, 3, 4):
        run = _solve(g, chains=chains)
        assert run[0]["n_chains"] == chains and run[4]["n_chains"] == chains and run[4]["n_border_nodes"] > 0
        _same_run(run, plain)
        assert run[5].max() < 1e-9 and plain[5].max() < 1e-9          # backward error of every linear solve
    O = load_oracle(g); so = O.solve(); qo, to = O.poses()
    s, q, t, sw = _solve(g, chains=2)[:4]
    assert len(s["iterations"]) == len(so["iterations"]) and abs(s["final_cost"] - so["final_cost"]) <= 1e-5 * so["final_cost"]
    assert np.abs(t - to).max() < 1e-5 and rot_angle_between(q, qo).max() < 1e-4 and np.array_equal(sw > 0.5, O.switches() > 0.5)


def test_default_plan_uses_two_chains_from_4096_nodes_on():
    p = problems.build_problem(2, n_nodes=6000, n_loop=600)
    g = _as_graph(p)
    auto = _solve(g)
    assert auto[0]["n_chains"] == 2 and auto[4]["n_chains"] == 2
    _same_run(auto, _solve(g, chains=1))
    small = _solve(_as_graph(problems.build_problem(3, n_nodes=1500, n_loop=300)))
    assert small[0]["n_chains"] == 1


def test_chains_fall_back_to_one_chain_when_the_cut_crosses_nothing_or_everything():
    # two disconnected halves: no edge crosses the middle cut -> nothing to split
    g = random_graph(400, 1, 0, seed=3)
    keep = ~((g["oc1"] >= 200) & (g["oc2"] < 200))
    for k in ("oc1", "oc2", "oq", "ot", "ow"):
        g[k] = g[k][keep]
    g["rn"] = np.array([0, 200], np.int32); g["rq"] = g["q"][[0, 200]].copy(); g["rt"] = g["t"][[0, 200]].copy(); g["rw"] = np.array([1.5, 1.5])
    a, b = _solve(g, chains=2), _solve(g, chains=1)
    assert a[0]["n_chains"] == 1
    _same_run(a, b)
    # loop closures across the whole trajectory: the separator would be most of the graph
    g = random_graph(300, 2, 260, seed=4)
    a, b = _solve(g, chains=2), _solve(g, chains=1)
    _same_run(a, b)


def test_constant_nodes_with_chains():
    g = random_graph(900, 3, 200, seed=7)
    runs = []
    for chains in (1, 2):
        S = load_pgs(g, chains=chains)
        S.set_constant_nodes(0, 40); S.set_constant_nodes(430, 40)      # a restored stretch at the start and one across the middle cut
        s = S.solve(); q, t = S.poses(); runs.append((s, q, t, S.switches())); S.close()
        assert np.array_equal(q[:40], g["q"][:40]) and np.array_equal(t[430:470], g["t"][430:470])
    _same_run(runs[1], runs[0])


def _ranks_on_one_gpu(g, world, **opt):
    group = "t-" + uuid.uuid4().hex
    out = [None] * world; errs = []

    def work(rank):
        try:
            S = load_pgs(g, **opt)
            S.dist_init_local(rank, world, group)
            s = S.solve(); q, t = S.poses()
            out[rank] = (s, q, t, S.switches(), S.dist_stats(), S.linear_backward_errors()); S.close()
        except Exception as ex:   # noqa: BLE001
            errs.append((rank, repr(ex)))
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [x.start() for x in th]; [x.join(timeout=600) for x in th]
    assert not errs and all(o is not None for o in out), errs
    return out


@pytest.mark.parametrize("world", [2, 3, 4, 5])
def test_ranks_over_the_local_transport_match_the_single_gpu_solve(world):
    """The multi-GPU algorithm with `world` ranks on one device.  world >= 3 is the case where a rank holds border
    nodes none of its own blocks touch (they came back at their initial guess)."""
    g = random_graph(1800, 3, 500, outlier_frac=0.1, seed=11)
    plain = _solve(g, chains=1)
    ranks = _ranks_on_one_gpu(g, world)
    for r in range(world):
        _same_run(ranks[r], plain)
        st = ranks[r][4]
        assert st["world"] == world and st["rank"] == r and st["n_border_nodes"] > 0 and st["n_collectives"] > 0
        assert ranks[r][5].max() < 1e-9
        # every rank ends with the complete solution, bit for bit the same
        assert np.array_equal(ranks[r][1], ranks[0][1]) and np.array_equal(ranks[r][2], ranks[0][2]) and np.array_equal(ranks[r][3], ranks[0][3])
    # a rank between the two free ends holds the separators at both of its ends, not the whole border
    if world >= 4:
        assert ranks[1][4]["n_local_border_nodes"] <= ranks[1][4]["n_border_nodes"]


def test_ranks_with_a_far_reaching_loop_edge():
    g = random_graph(1200, 3, 200, seed=13)
    g["la"] = np.append(g["la"], 1199).astype(np.int32); g["lb"] = np.append(g["lb"], 0).astype(np.int32)
    g["lq"] = np.vstack([g["lq"], [0, 0, 0, 1.0]]); g["lt"] = np.vstack([g["lt"], g["gt_t"][1199] - g["gt_t"][0]]); g["lw"] = np.append(g["lw"], 1.0)
    plain = _solve(g, chains=1)
    for r in _ranks_on_one_gpu(g, 3):
        _same_run(r, plain)
    _same_run(_solve(g, chains=2), plain)
