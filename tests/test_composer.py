"""Composer pose assembly (reference src/Composer.cpp:10-292).
CPU: known-answer cases pin the oracle restatement (oracle/composer.py).  GPU: the device gather behind
include/pgs_compose.h and the facade's Composer against the oracle on multi-world graphs with dead zones, before
any solve, after a solve, and with keyframes that arrived after the solve."""
import numpy as np
import pytest

from oracle import composer as ocomp
from oracle import frontend, pgo
from solve_keyframe_pose_graph_b200 import capi, facade, synth

TOL = 1e-9   # fp64; the device uses the rigid inverse, the reference / oracle the generic 4x4 inverse


def _rand_pose(rng, scale=5.0):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return pgo.pose_to_mat4(q, rng.normal(size=3) * scale)


def _manager(g):
    M = frontend.Manager(); M.ingest(g)
    return M


def test_nothing_solved_gives_odometry_poses():
    g = synth.generate_config(4, n_nodes=40, n_interworld=0)
    M = _manager(g)
    T, wid, jmb = ocomp.assemble(M, [], 0)
    assert np.allclose(T, np.array(M.poses), atol=0)          # solvedUntil == 0: w_TM_i = manager pose (Composer.cpp:128-132)
    assert sorted(set(wid)) == [-3, -2, -1, 0, 1, 2, 3] and sum(len(v) for v in jmb.values()) == len(T)


def test_solved_poses_equal_to_odometry_reproduce_odometry():
    # if the optimiser returns the odometry poses, carrying them forward through odometry is the identity
    g = synth.generate_config(2, n_nodes=120, n_loop=4)
    M = _manager(g)
    su = 70
    T, wid, _ = ocomp.assemble(M, [P.copy() for P in M.poses[:su + 1]], su)
    assert np.allclose(T, np.array(M.poses), atol=1e-9) and np.all(wid == 0)


def test_rigid_correction_is_carried_forward_and_through_dead_zones():
    # three worlds of 30 keyframes with dead zones of 5 in between.  The optimiser moved world 0 by a rigid D.
    # (three worlds, not two: with exactly one kidnap the reference's which_world_is_this puts the keyframe AT the
    # kidnap stamp into the dead zone, NodeDataManager.cpp:1141 vs :1160 — restated literally by the oracle)
    g = synth.generate_config(4, n_nodes=30, n_worlds=3, n_interworld=0)
    M = _manager(g)
    rng = np.random.default_rng(5)
    D = _rand_pose(rng)
    su = 19
    slam = [D @ P for P in M.poses[:su + 1]]
    T, wid, _ = ocomp.assemble(M, slam, su)
    w0 = np.where(wid == 0)[0]; dz = np.where(wid == -1)[0]; w1 = np.where(wid == 1)[0]
    assert len(w0) == 30 and len(dz) == 5 and len(w1) == 30 and w0.max() > su
    for i in w0:   # solved ones are D*M_i; unsolved ones of the same world: D*M_su*inv(M_su)*M_i = D*M_i (:155-165)
        assert np.allclose(T[i], D @ M.poses[i], atol=1e-9)
    for i in dz:   # dead zone hangs off the last assembled pose of world 0 (:140-146)
        assert np.allclose(T[i], D @ M.poses[i], atol=1e-9)
    for i in w1:   # another world, not yet solved: raw odometry (:137-139)
        assert np.allclose(T[i], M.poses[i], atol=0)


def _facade_inputs(F, M):
    """Arrays of include/pgs_compose.h from the oracle manager + the facade's optimised poses."""
    n = len(M.poses)
    q, t = F.poses()
    wid = np.array([M.which_world_is_this(s) for s in M.stamps], np.int32)
    nw = M.n_worlds()
    end = np.array([M.nodeidx_of_world_i_ended(w) for w in range(nw)], np.int32)
    sid = np.array([M.worlds.find_setID_of_world_i(w) for w in range(nw)], np.int32)
    ex = np.zeros(nw, np.uint8); wt = np.tile(np.eye(4), (nw, 1, 1))
    for w in range(nw):
        if sid[w] != w and M.worlds.is_exist(int(sid[w]), w):
            ex[w] = 1; wt[w] = M.worlds.getPoseBetweenWorlds(int(sid[w]), w)
    return np.array(M.poses), wid, q, t, end, sid, ex, wt


@pytest.mark.gpu
@pytest.mark.parametrize("n_nodes,n_worlds,n_inter", [(150, 4, 24), (400, 3, 12)])
def test_device_compose_matches_oracle_through_a_session(n_nodes, n_worlds, n_inter):
    g = synth.generate_config(4, n_nodes=n_nodes, n_worlds=n_worlds, n_interworld=n_inter)
    cut = g["N"] - 37                                  # the last keyframes arrive after the solve
    F = facade.Facade(odom_fanout=3)
    M = frontend.Manager()
    early = {k: (v[:cut] if k in ("stamps", "q", "t") else v) for k, v in g.items()}; early["N"] = cut
    keep = (g["la"] < cut) & (g["lb"] < cut)
    for k in ("la", "lb", "lq", "lt", "lw"):
        early[k] = g[k][keep]
    F.ingest(early); M.ingest(early)

    def check(slam_poses, su):
        T, wid = F.compose()
        To, wo, _ = ocomp.assemble(M, slam_poses, su)
        assert np.array_equal(wid, wo)
        assert np.abs(T - To).max() < TOL * max(1.0, np.abs(To).max())
        # the raw C-ABI agrees with the facade path
        mgr_T, w_, q, t, end, sid, ex, wt = _facade_inputs(F, M)
        ns = F.n_nodes()
        Tr, ms_k, ms_t = capi.compose_poses(mgr_T, w_, q[:ns], t[:ns], su, M.which_world_is_this(M.stamps[su]), end, sid, ex, wt)
        assert np.abs(Tr - To).max() < TOL * max(1.0, np.abs(To).max()) and ms_k > 0 and ms_t >= ms_k
        idx, Tl, st = F.last_known_camerapose()
        assert idx == len(M.poses) - 1 and st == M.stamps[-1] and np.allclose(Tl, To[-1], atol=TOL * max(1.0, np.abs(To).max()))

    check([], 0)                                       # before the first solve
    assert F.solve_once()
    R = frontend.ReferenceFrontEnd(M, odom_fanout=3); R.trigger(solve=False)   # merges the worlds exactly like the facade's trigger
    q, t = F.poses()
    slam = [pgo.pose_to_mat4(q[i], t[i]) for i in range(len(q))]
    check(slam, F.solved_until())
    # keyframes (and a dead zone) that arrive after the solve are carried forward through odometry
    F.add_nodes(g["stamps"][cut:], g["q"][cut:], g["t"][cut:])
    for i in range(cut, g["N"]):
        M.add_node(g["stamps"][i], g["q"][i], g["t"][i])
    check(slam, F.solved_until())
    F.close()
