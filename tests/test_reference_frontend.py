"""The front-end rules against THE REFERENCE'S OWN FRONT END (SURVEY §8a rows "graph construction", "initial guesses",
"world / set bookkeeping", "variable store"; §8f-1 incremental triggers).

oracle/_ref/libref_frontend.so is the reference's src/NodeDataManager.cpp, src/Worlds.cpp, src/PoseGraphSLAM.cpp,
src/utils/PoseManipUtils.cpp and src/utils/RawFileIO.cpp compiled UNMODIFIED, from where they lie, over oracle/shim/
(stand-ins for the Eigen / Ceres / roscpp / message / OpenCV names they use; wrapper oracle/ref_frontend_capi.cpp).  The
ROS callbacks receive messages built from the same arrays the product's facade ingests; the reference's solver thread
runs PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_optimize6DOF() itself, single-stepped through a gate in the
ros::Rate stand-in; the stand-in ceres::Problem records what the reference builds and ceres::Solve snapshots it without
minimising.  The product's facade (dry run: same rule, no device) and the oracle's Python front-end must have built, after
every wake-up: the same residual blocks between the same keyframes with the same observations and weights, the same
regularisers, the same initial guess for every keyframe, the same solvedUntil and the same world / set state.
Built only where the reference tree exists; the tests skip without the library."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import frontend, pgo
from solve_keyframe_pose_graph_b200 import facade, synth

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_frontend.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_frontend.so not built (needs /root/reference; make -C oracle)")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


class ReferenceNode:
    """The reference's NodeDataManager + PoseGraphSLAM behind oracle/ref_frontend_capi.cpp."""

    def __init__(self):
        # NodeDataManager::camera_pose_callback keeps a function-local `static bool` for "first keyframe ever"
        # (src/NodeDataManager.cpp:25): one manager per loaded image.  Every instance here loads its own copy of the library.
        import shutil
        import tempfile
        self._dir = tempfile.mkdtemp(prefix="refslam_")
        private_copy = os.path.join(self._dir, "libref_frontend.so")
        shutil.copy(REF_SO, private_copy)
        L = self.L = C.CDLL(private_copy)
        L.refslam_create.restype = C.c_void_p
        for f, extra in dict(refslam_destroy=[], refslam_wakeup=[], refslam_n_blocks=[], refslam_unknown_blocks=[], refslam_n_vars=[], refslam_n_switches=[],
                             refslam_solved_until=[], refslam_status=[], refslam_n_nodes=[], refslam_n_edges=[], refslam_n_worlds=[],
                             refslam_world_setid=[C.c_int], refslam_world_start=[C.c_int], refslam_world_end=[C.c_int], refslam_which_world=[C.c_longlong],
                             refslam_add_node=[C.c_longlong, dp, dp], refslam_add_loop_edge=[C.c_longlong, C.c_longlong, dp, dp, C.c_float],
                             refslam_kidnap=[C.c_longlong, C.c_int], refslam_get_blocks=[ip, ip, ip, ip, dp, dp], refslam_get_vars=[dp, dp, dp, ip],
                             refslam_pose_between_worlds=[C.c_int, C.c_int, dp], refslam_get_node_pose=[C.c_int, dp], refslam_set_perturb=[C.c_double]).items():
            getattr(L, f).argtypes = [C.c_void_p] + extra
        self.h = L.refslam_create()
        assert self.h, "one reference instance at a time"

    def close(self):
        if self.h:
            self.L.refslam_destroy(self.h); self.h = None
            import shutil
            shutil.rmtree(self._dir, ignore_errors=True)

    def add_nodes(self, stamps, q, t):
        for s, qq, tt in zip(stamps, np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(t, dtype=np.float64)):
            self.L.refslam_add_node(self.h, int(s), qq.ctypes.data_as(dp), tt.ctypes.data_as(dp))

    def add_loop_edges(self, stamps, a, b, q, t, w):
        for aa, bb, qq, tt, ww in zip(a, b, np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(t, dtype=np.float64), w):
            self.L.refslam_add_loop_edge(self.h, int(stamps[aa]), int(stamps[bb]), qq.ctypes.data_as(dp), tt.ctypes.data_as(dp), float(ww))

    def kidnap(self, stamp, kidnapped):
        self.L.refslam_kidnap(self.h, int(stamp), int(kidnapped))

    def wakeup(self):
        rc = self.L.refslam_wakeup(self.h)
        assert rc >= 0, "the reference's solver thread did not come back to its sleep within 120 s"
        return rc == 1

    def blocks(self):
        n = self.L.refslam_n_blocks(self.h)
        assert self.L.refslam_unknown_blocks(self.h) == 0
        ty = np.zeros(n, np.int32); c1 = np.zeros(n, np.int32); c2 = np.zeros(n, np.int32); sw = np.zeros(n, np.int32); obs = np.zeros((n, 4, 4)); w = np.zeros(n)
        self.L.refslam_get_blocks(self.h, ty.ctypes.data_as(ip), c1.ctypes.data_as(ip), c2.ctypes.data_as(ip), sw.ctypes.data_as(ip), obs.ctypes.data_as(dp), w.ctypes.data_as(dp))
        return dict(type=ty, c1=c1, c2=c2, sw=sw, obs=obs, w=w)

    def variables(self):
        n, m = self.L.refslam_n_vars(self.h), self.L.refslam_n_switches(self.h)
        q = np.zeros((n, 4)); t = np.zeros((n, 3)); s = np.zeros(max(m, 1)); c = np.zeros(n, np.int32)
        self.L.refslam_get_vars(self.h, q.ctypes.data_as(dp), t.ctypes.data_as(dp), s.ctypes.data_as(dp), c.ctypes.data_as(ip))
        return q, t, s[:m], c

    def pose_between_worlds(self, m, n):
        T = np.zeros((4, 4))
        return T if self.L.refslam_pose_between_worlds(self.h, m, n, T.ctypes.data_as(dp)) else None


def mats(q, t):
    return np.array([pgo.pose_to_mat4(qq, tt) for qq, tt in zip(q, t)]).reshape(-1, 4, 4)


def same_quats(a, b):
    return len(a) == len(b) and (len(a) == 0 or np.all(np.abs(np.sum(a * b, axis=1)) > 1 - 1e-12))


def compare(R, F, P, fanout):
    """R: the reference (real code), F: the product's facade (dry run), P: the oracle's Python front-end — after a wake-up that triggered."""
    B = R.blocks()
    od, lo, rg = B["type"] == 0, B["type"] == 1, B["type"] == 2
    # ---- odometry blocks: SixDOFError::Create(u_M_umf, odom_edge_weight) on (u, u-f)   [PoseGraphSLAM.cpp:1570-1639]
    if F is not None:
        o = F.alternative_terms(0)                                    # the facade's odometry blocks with their observations
        assert np.array_equal(B["c1"][od], o["c1"]) and np.array_equal(B["c2"][od], o["c2"])
        assert np.allclose(B["w"][od], o["weight"], rtol=1e-9, atol=0)
        assert np.allclose(B["obs"][od], mats(o["obs_rot"], o["obs_t"]), rtol=0, atol=1e-9)
    assert np.allclose(B["obs"][od], mats([x[2] for x in P.odom], [x[3] for x in P.odom]), rtol=0, atol=1e-9)
    assert [(x[0], x[1]) for x in P.odom] == list(zip(B["c1"][od].tolist(), B["c2"][od].tolist()))
    assert np.allclose([x[4] for x in P.odom], B["w"][od], rtol=1e-9, atol=0)
    # ---- loop blocks: SixDOFErrorWithSwitchingConstraints::Create(bTa, weight) on (second, first, switch e)   [:1381-1559]
    if F is not None:
        l = F.alternative_terms(1)
        assert np.array_equal(B["c1"][lo], l["c1"]) and np.array_equal(B["c2"][lo], l["c2"])
        assert np.allclose(B["obs"][lo], mats(l["obs_rot"], l["obs_t"]), rtol=0, atol=1e-9) and np.allclose(B["w"][lo], l["weight"], rtol=1e-6)
    assert np.allclose(B["obs"][lo], mats([x[3] for x in P.loops], [x[4] for x in P.loops]), rtol=0, atol=1e-9)
    assert [(x[2], x[1], x[0]) for x in P.loops] == list(zip(B["c1"][lo].tolist(), B["c2"][lo].tolist(), B["sw"][lo].tolist()))
    # ---- regularisers: one per set-root world, anchored at the CURRENT estimate of its first keyframe   [:1801-1850]
    if F is not None:
        r = F.reg_terms()
        assert np.array_equal(B["c1"][rg], r["node"]) and np.allclose(B["w"][rg], r["w"], rtol=1e-12)
        assert np.allclose(B["obs"][rg], mats(r["q"], r["t"]), rtol=0, atol=1e-9)
    assert [x[0] for x in P.regs] == B["c1"][rg].tolist() and np.allclose([x[3] for x in P.regs], B["w"][rg], rtol=1e-12)
    assert np.allclose(B["obs"][rg], mats([x[1] for x in P.regs], [x[2] for x in P.regs]), rtol=0, atol=1e-9)
    # ---- optimisation variables = the initial guesses of this wake-up   [:226-361, 1649-1793]
    q, t, s, const = R.variables()
    if F is not None:
        fq, ft = F.poses()
        assert len(t) == len(ft) and np.allclose(t, ft, rtol=0, atol=1e-8) and same_quats(q, fq)
        assert np.all(s == 0.99)                                      # switches start at 0.99 (:353) and nothing here moves them
    assert np.allclose(t, np.array(P.opt_t).reshape(-1, 3), rtol=0, atol=1e-8) and same_quats(q, np.array(P.opt_q).reshape(-1, 4))
    assert np.allclose(s, P.opt_s, rtol=0, atol=1e-15) and not const.any()   # nothing is marked constant in the live path
    # ---- bookkeeping after the wake-up
    assert R.L.refslam_solved_until(R.h) == P.solved_until and (F is None or F.solved_until() == P.solved_until)
    assert R.L.refslam_n_worlds(R.h) == P.m.n_worlds() and (F is None or F.n_worlds() == P.m.n_worlds())
    for w in range(P.m.n_worlds()):
        assert R.L.refslam_world_setid(R.h, w) == P.m.worlds.find_setID_of_world_i(w) and (F is None or F.world_setid(w) == P.m.worlds.find_setID_of_world_i(w))
        assert R.L.refslam_world_start(R.h, w) == P.m.nodeidx_of_world_i_started(w) and R.L.refslam_world_end(R.h, w) == P.m.nodeidx_of_world_i_ended(w)
        assert F is None or (F.world_start(w), F.world_end(w)) == (P.m.nodeidx_of_world_i_started(w), P.m.nodeidx_of_world_i_ended(w))
    return B


def test_single_world_session_in_three_wakeups_matches_the_reference_front_end():
    g = synth.generate_config(2, n_nodes=240, n_loop=36)
    order = np.argsort(np.maximum(g["la"], g["lb"]), kind="stable")
    R = ReferenceNode(); F = facade.Facade(odom_fanout=5, dry_run=True); M = frontend.Manager(); P = frontend.ReferenceFrontEnd(M, odom_fanout=5)
    try:
        epos = 0; triggered = 0
        for lo in range(0, 240, 80):
            sl = slice(lo, lo + 80)
            R.add_nodes(g["stamps"][sl], g["q"][sl], g["t"][sl]); F.add_nodes(g["stamps"][sl], g["q"][sl], g["t"][sl])
            for i in range(lo, lo + 80):
                M.add_node(int(g["stamps"][i]), g["q"][i], g["t"][i])
            take = []
            while epos < len(order) and max(g["la"][order[epos]], g["lb"][order[epos]]) < lo + 80:
                take.append(order[epos]); epos += 1
            take = np.array(take, dtype=int)
            if len(take):
                R.add_loop_edges(g["stamps"], g["la"][take], g["lb"][take], g["lq"][take], g["lt"][take], g["lw"][take])
                F.add_loop_edges(g["la"][take], g["lb"][take], g["lq"][take], g["lt"][take], g["lw"][take])
                for e in take:
                    M.add_loop_edge(int(g["la"][e]), int(g["lb"][e]), g["lq"][e], g["lt"][e], float(g["lw"][e]))
            fired = R.wakeup()
            assert fired == F.solve_once() == (len(take) > 0)
            if fired:
                P.trigger(solve=False); triggered += 1
                B = compare(R, F, P, 5)
                assert (B["type"] == 0).sum() == sum(min(5, u) for u in range(lo + 80)) and (B["type"] == 1).sum() == epos and (B["type"] == 2).sum() == 1
            assert not R.wakeup() and not F.solve_once()                      # nothing new: the loop goes back to sleep (:1306-1312)
        assert triggered >= 2 and epos == len(order)
    finally:
        R.close(); F.close()


def test_two_world_session_with_a_kidnap_matches_the_reference_front_end():
    """World 0, a kidnap, world 1, then loop edges from world 1 back into world 0: the first of them fixes the relative
    pose of the worlds from odometry (:1459-1464), the sets merge, and every keyframe of world 1 is initialised in the
    frame of the set root."""
    rng = np.random.default_rng(9)
    g = synth.generate_config(4, n_nodes=60, n_interworld=10)
    stamps, k0, k1 = g["stamps"], g["k0"], g["k1"]
    w0 = np.nonzero(stamps <= k0[0])[0]; w1 = np.nonzero((stamps > k1[0]) & (stamps <= k0[1]))[0]
    dead = np.nonzero((stamps > k0[0]) & (stamps <= k1[0]))[0]
    keep = np.r_[w0, dead, w1]
    assert np.array_equal(keep, np.arange(len(keep)))
    n = len(keep)
    R = ReferenceNode(); F = facade.Facade(odom_fanout=5, dry_run=True); M = frontend.Manager(); P = frontend.ReferenceFrontEnd(M, odom_fanout=5)   # the reference hard-codes f = 1..5 (:1577)

    def nodes(idx):
        R.add_nodes(stamps[idx], g["q"][idx], g["t"][idx]); F.add_nodes(stamps[idx], g["q"][idx], g["t"][idx])
        for i in idx:
            M.add_node(int(stamps[i]), g["q"][i], g["t"][i])

    def loops(pairs):
        for a, b in pairs:                                                    # observation: noisy relative pose of the odometry frames, any value will do
            T = pgo.inv4(pgo.pose_to_mat4(g["q"][b], g["t"][b])) @ pgo.pose_to_mat4(g["q"][a], g["t"][a])
            q, t = pgo.mat4_to_pose(T); t = t + rng.normal(size=3) * 0.05
            R.add_loop_edges(stamps, [a], [b], [q], [t], [1.0]); F.add_loop_edges([a], [b], [q], [t], [1.0]); M.add_loop_edge(a, b, q, t, 1.0)

    try:
        # wake-up 1: world 0 with two loop edges inside it
        nodes(w0); loops([(int(w0[-5]), int(w0[3])), (int(w0[-12]), int(w0[8]))])
        assert R.wakeup() and F.solve_once(); P.trigger(solve=False)
        compare(R, F, P, 5)
        # kidnapped: keyframes of the dead zone arrive, a loop edge touching them is ignored, and nothing may trigger
        for X in (R, F, M):
            (X.kidnap if X is R else X.kidnap_indicator)(int(k0[0]), 1)
        nodes(dead)
        assert not R.wakeup() and not F.solve_once()
        for X in (R, F, M):
            (X.kidnap if X is R else X.kidnap_indicator)(int(k1[0]), 0)
        # wake-up 2: world 1 and three loop edges back into world 0
        nodes(w1); loops([(int(w1[10]), int(w0[20])), (int(w1[30]), int(w0[40])), (int(w1[-1]), int(w1[5]))])
        assert [R.L.refslam_which_world(R.h, int(s)) for s in stamps[:n]] == [F.which_world(int(s)) for s in stamps[:n]] == [M.which_world_is_this(int(s)) for s in stamps[:n]]
        assert R.wakeup() and F.solve_once(); P.trigger(solve=False)
        B = compare(R, F, P, 5)
        assert F.n_worlds() == 2 and F.world_setid(1) == F.world_setid(0) == 0
        T = R.pose_between_worlds(0, 1)
        assert T is not None and np.allclose(T, F.pose_between_worlds(0, 1), rtol=0, atol=1e-9) and np.allclose(T, M.worlds.getPoseBetweenWorlds(0, 1), rtol=0, atol=1e-9)
        assert (B["type"] == 2).sum() == 1 and B["c1"][B["type"] == 2][0] == 0                # one set, one regulariser, on keyframe 0
        no_dead = ~np.isin(B["c1"], dead) & ~np.isin(B["c2"], dead)
        assert no_dead[B["type"] != 2].all()                                                  # no block touches a dead-zone keyframe
    finally:
        R.close(); F.close()


def stand_in_solve(P, amplitude, k):
    """What oracle/ref_frontend_capi.cpp does to the reference's variables in place of ceres::Solve (same closed forms)."""
    used, sused = set(), set()
    for x in P.odom: used.update((x[0], x[1]))
    for x in P.loops: used.update((x[1], x[2])); sused.add(x[0])
    for x in P.regs: used.add(x[0])
    a = amplitude
    for i in sorted(used):
        e = np.array([a * 0.1 * np.sin(i + k), a * 0.1 * np.cos(2 * i + k), a * 0.1 * np.sin(3 * i + 2 * k)])
        P.opt_q[i] = pgo.quat_plus(np.asarray(P.opt_q[i], dtype=float), e)
        P.opt_t[i] = np.asarray(P.opt_t[i], dtype=float) + a * np.array([np.cos(i + k), np.sin(2 * i + k), np.cos(3 * i + k)])
    for e in sorted(sused):
        P.opt_s[e] = 0.99 - 0.01 * ((7 * e + k) % 50)


def test_warm_start_rules_after_the_state_has_moved_match_the_reference_front_end():
    """The rules that only show once a solve has MOVED the variables: keyframes that arrive after a solve are carried
    forward from the last solved pose through odometry (:1745-1760), a world that was solved on its own and is then merged
    into another set has its already-solved keyframes re-based into the new set root's frame (:1700-1730), regularisers
    are re-anchored at the current estimates, switch values persist.  ceres::Solve is replaced on the reference's side by
    a known index-dependent displacement of every variable in the problem; the Python front-end gets the same displacement.
    (The facade has no entry point to overwrite its variables; its warm-start rules are compared with the Python front-end
    under real solves in tests/test_incremental_gpu.py.)"""
    rng = np.random.default_rng(10)
    g = synth.generate_config(4, n_nodes=60, n_interworld=10)
    stamps, k0, k1 = g["stamps"], g["k0"], g["k1"]
    w0 = np.nonzero(stamps <= k0[0])[0]; dead = np.nonzero((stamps > k0[0]) & (stamps <= k1[0]))[0]; w1 = np.nonzero((stamps > k1[0]) & (stamps <= k0[1]))[0]
    R = ReferenceNode(); M = frontend.Manager(); P = frontend.ReferenceFrontEnd(M, odom_fanout=5)
    R.L.refslam_set_perturb(R.h, 0.3)

    def nodes(idx):
        R.add_nodes(stamps[idx], g["q"][idx], g["t"][idx])
        for i in idx:
            M.add_node(int(stamps[i]), g["q"][i], g["t"][i])

    def loops(pairs):
        for a, b in pairs:
            T = pgo.inv4(pgo.pose_to_mat4(g["q"][b], g["t"][b])) @ pgo.pose_to_mat4(g["q"][a], g["t"][a])
            q, t = pgo.mat4_to_pose(T); t = t + rng.normal(size=3) * 0.05
            R.add_loop_edges(stamps, [a], [b], [q], [t], [1.0]); M.add_loop_edge(a, b, q, t, 1.0)

    def wake(k):
        assert R.wakeup(); P.trigger(solve=False)
        B = compare(R, None, P, 5)                                            # the snapshot is taken BEFORE the stand-in solve moves anything
        stand_in_solve(P, 0.3, k)
        return B

    try:
        nodes(w0[:40]); loops([(int(w0[35]), int(w0[3]))]); wake(1)           # 1: first solve of world 0
        nodes(w0[40:]); loops([(int(w0[-2]), int(w0[10]))])                   # 2: keyframes that arrived after it are carried forward
        B = wake(2)
        assert (B["type"] == 2).sum() == 1
        R.kidnap(int(k0[0]), 1); M.kidnap_indicator(int(k0[0]), 1); nodes(dead); assert not R.wakeup()
        R.kidnap(int(k1[0]), 0); M.kidnap_indicator(int(k1[0]), 0)
        nodes(w1[:40]); loops([(int(w1[30]), int(w1[4]))])                    # 3: world 1 solved on its own: a second set root, a second regulariser
        B = wake(3)
        assert (B["type"] == 2).sum() == 2 and M.worlds.find_setID_of_world_i(1) == 1
        nodes(w1[40:]); loops([(int(w1[45]), int(w0[20]))])                   # 4: the first edge into world 0 merges the sets: solved keyframes of world 1 are re-based
        B = wake(4)
        assert (B["type"] == 2).sum() == 1 and M.worlds.find_setID_of_world_i(1) == 0
        T = R.pose_between_worlds(0, 1)
        assert T is not None and np.allclose(T, M.worlds.getPoseBetweenWorlds(0, 1), rtol=0, atol=1e-9)
        loops([(int(w1[50]), int(w0[50])), (int(w1[55]), int(w1[12]))]); wake(5)   # 5: further edges on the merged set
    finally:
        R.close()


def _two_world_session(R, F, rng):
    """World 0, a kidnap with dead-zone keyframes, world 1, loop edges inside and across the worlds; one wake-up at the end."""
    g = synth.generate_config(4, n_nodes=40, n_interworld=6)
    stamps, k0, k1 = g["stamps"], g["k0"], g["k1"]
    w0 = np.nonzero(stamps <= k0[0])[0]; dead = np.nonzero((stamps > k0[0]) & (stamps <= k1[0]))[0]; w1 = np.nonzero((stamps > k1[0]) & (stamps <= k0[1]))[0]
    for X in (R, F):
        X.add_nodes(stamps[w0], g["q"][w0], g["t"][w0])
        (X.kidnap if X is R else X.kidnap_indicator)(int(k0[0]), 1)
        X.add_nodes(stamps[dead], g["q"][dead], g["t"][dead])
        (X.kidnap if X is R else X.kidnap_indicator)(int(k1[0]), 0)
        X.add_nodes(stamps[w1], g["q"][w1], g["t"][w1])
    for a, b in [(int(w0[30]), int(w0[2])), (int(w1[10]), int(w0[20])), (int(w1[35]), int(w1[4])), (int(w1[20]), int(w0[5]))]:
        T = pgo.inv4(pgo.pose_to_mat4(g["q"][b], g["t"][b])) @ pgo.pose_to_mat4(g["q"][a], g["t"][a])
        q, t = pgo.mat4_to_pose(T); t = t + rng.normal(size=3) * 0.05
        R.add_loop_edges(stamps, [a], [b], [q], [t], [1.0]); F.add_loop_edges([a], [b], [q], [t], [1.0])
    assert R.wakeup() and F.solve_once()
    return g, len(w0) + len(dead) + len(w1)


def test_state_files_equal_the_files_the_reference_code_writes_and_its_loader_reads_ours(tmp_path):
    """NodeDataManager::saveAsJSON, PoseGraphSLAM::saveAsJSON and Worlds::saveStateToDisk of the reference (real code) against
    the product's writers for the same two-world session: log_posegraph.json and log_optimized_poses.json byte for byte
    (the product adds two exact-nanosecond fields per kidnap, which the reference's loader ignores), the WorldsData object
    value for value.  Then NodeDataManager::loadFromJSON and Worlds::loadStateFromDisk of the reference read the PRODUCT's
    files into a fresh reference instance, which must end up in the same state."""
    import json
    rng = np.random.default_rng(12)
    R = ReferenceNode(); F = facade.Facade(odom_fanout=5, dry_run=True)
    R.L.refslam_save_json.argtypes = [C.c_void_p, C.c_char_p]
    dR, dF = tmp_path / "ref", tmp_path / "ours"; dR.mkdir(); dF.mkdir()
    try:
        g, n = _two_world_session(R, F, rng)
        assert R.L.refslam_save_json(R.h, str(dR).encode()) == 7
        F.save_json(dF)
        import re
        number = re.compile(r"-?\d+\.?\d*(?:[eE][+-]?\d+)?")
        for name in ("log_posegraph.json", "log_optimized_poses.json"):
            ref_lines = open(dR / name).read().split("\n")
            our_lines = [l for l in open(dF / name).read().split("\n") if '"stampNSec_started"' not in l and '"stampNSec_ended"' not in l]
            assert len(ref_lines) == len(our_lines) and ref_lines[-1] == our_lines[-1] == "", name          # both end with one newline
            identical = 0
            for i, (a, b) in enumerate(zip(ref_lines, our_lines)):
                if a == b:
                    identical += 1; continue
                # a line may differ only in the last digits of matrix entries: the arithmetic under the reference's code here is
                # the stand-in's, not Eigen's, so products are not bit-equal; layout, keys, order and every other character are
                assert number.sub("#", a) == number.sub("#", b), (name, i, a, b)
                va, vb = [float(x) for x in number.findall(a)], [float(x) for x in number.findall(b)]
                assert np.allclose(va, vb, rtol=1e-12, atol=1e-9), (name, i, a, b)
            assert identical >= 0.9 * len(ref_lines), (name, identical, len(ref_lines))
        wr = json.load(open(dR / "worlds.json")); wo = json.load(open(dF / "solved_posegraph.json"))["WorldsData"]
        for a, b in zip(json.dumps(wr, indent=1, sort_keys=True).split("\n"), json.dumps(wo, indent=1, sort_keys=True).split("\n")):
            assert number.sub("#", a) == number.sub("#", b), (a, b)                 # the matrix string differs in last digits only (see above)
            assert np.allclose([float(x) for x in number.findall(a)], [float(x) for x in number.findall(b)], rtol=1e-12, atol=1e-9), (a, b)
        assert wr["disjoint_set"] == wo["disjoint_set"] and wr["vec_world_starts"] == wo["vec_world_starts"] and wr["vec_world_ends"] == wo["vec_world_ends"]
        assert wr["disjoint_set"]["log_string"] == "add_element:0;add_element:1;union_sets:1,0;" and len(wr["rel_pose_between_worlds__wb_T_wa"]) == 1
    finally:
        R.close()
    # ---- the reference's loaders on the product's files
    R2 = ReferenceNode()
    for f, a in dict(refslam_load_posegraph_json=[C.c_char_p], refslam_load_worlds_json=[C.c_char_p], refslam_node_stamp=[C.c_int], refslam_manager_node_pose=[C.c_int, dp],
                     refslam_edge=[C.c_int, ip, ip, dp, dp], refslam_n_kidnaps=[]).items():
        getattr(R2.L, f).argtypes = [C.c_void_p] + a
    R2.L.refslam_node_stamp.restype = C.c_longlong
    try:
        json.dump(json.load(open(dF / "solved_posegraph.json"))["WorldsData"], open(tmp_path / "worldsdata.json", "w"))
        assert R2.L.refslam_load_worlds_json(R2.h, str(tmp_path / "worldsdata.json").encode()) == 1
        assert R2.L.refslam_load_posegraph_json(R2.h, str(dF).encode()) == 1      # keyframes and loop edges; kidnap stamps are not part of what this loader restores
        assert R2.L.refslam_n_nodes(R2.h) == n == F.n_keyframes() and R2.L.refslam_n_edges(R2.h) == 4
        T = np.zeros((4, 4)); a = C.c_int(); b = C.c_int(); w = C.c_double()
        for i in range(n):
            assert abs(R2.L.refslam_node_stamp(R2.h, i) - int(g["stamps"][i])) < 1000    # stamps travel as seconds in a double
            R2.L.refslam_manager_node_pose(R2.h, i, T.ctypes.data_as(dp))
            assert np.allclose(T, pgo.pose_to_mat4(g["q"][i], g["t"][i]), rtol=0, atol=1e-12)
        for e in range(4):
            R2.L.refslam_edge(R2.h, e, C.byref(a), C.byref(b), T.ctypes.data_as(dp), C.byref(w))
            assert w.value == 1.0 and 0 <= b.value < a.value < n
        assert [R2.L.refslam_world_setid(R2.h, k) for k in range(2)] == [F.world_setid(k) for k in range(2)] == [0, 0]
        P01 = R2.pose_between_worlds(0, 1)
        assert P01 is not None and np.allclose(P01, F.pose_between_worlds(0, 1), rtol=1e-14, atol=1e-12)
    finally:
        R2.close(); F.close()


def test_restoring_a_saved_session_matches_the_reference_restore_pipeline(tmp_path):
    """Composer::loadStateFromDisk — what the reference node does for its `loadStateFromDisk` parameter — is four calls of
    the reference's own functions (Worlds::loadStateFromDisk, NodeDataManager::load_kidnap_data_from_json,
    NodeDataManager::load_solved_posegraph_data_from_json, PoseGraphSLAM::load_state).  They and the product's
    pgs_facade_load_state_from_disk restore the same solved_posegraph.json; afterwards both sessions go on with new keyframes
    and a loop edge into the restored (constant) part, and must build the same problem."""
    import json
    rng = np.random.default_rng(13)
    R = ReferenceNode(); F = facade.Facade(odom_fanout=5, dry_run=True)
    try:
        g, n = _two_world_session(R, F, rng)
        F.save_json(tmp_path)                                                # WorldsData + KidnapTimestamps; SolvedPoseGraph is filled below
        q, t = F.poses()                                                     # every keyframe in the frame of its set root (dead-zone ones: placeholders)
        world = [F.which_world(int(s)) for s in g["stamps"][:n]]
    finally:
        R.close(); F.close()
    J = json.load(open(tmp_path / "solved_posegraph.json"))
    assert J["SolvedPoseGraph"] in ([], None)                                # no Composer pass without a device: write what it would have written
    J["SolvedPoseGraph"] = []
    for i in range(n):
        T = pgo.pose_to_mat4(q[i], t[i]) if world[i] >= 0 else pgo.pose_to_mat4(g["q"][i], g["t"][i])
        J["SolvedPoseGraph"].append(dict(seq=i, stampNSec=int(g["stamps"][i]), worldID=world[i], setID_of_worldID=0 if world[i] >= 0 else -1,
                                         w_T_c=dict(rows=4, cols=4, data=facade.io_mat_to_string(T, solved_layout=True), data_pretty=facade.io_prettyprint(T))))
    d = tmp_path / "restore"; d.mkdir()
    json.dump(J, open(d / "solved_posegraph.json", "w"), indent=4)

    R = ReferenceNode(); F = facade.Facade(odom_fanout=5, dry_run=True)
    for f, a in dict(refslam_load_state_from_disk=[C.c_char_p], refslam_slam_n_nodes=[], refslam_kidnap_status=[], refslam_n_kidnaps=[], refslam_node_stamp=[C.c_int],
                     refslam_manager_node_pose=[C.c_int, dp]).items():
        getattr(R.L, f).argtypes = [C.c_void_p] + a
    R.L.refslam_node_stamp.restype = C.c_longlong
    try:
        assert R.L.refslam_load_state_from_disk(R.h, str(d).encode()) == 0
        F.load_state_from_disk(d)
        assert R.L.refslam_n_nodes(R.h) == F.n_keyframes() == n and R.L.refslam_slam_n_nodes(R.h) == F.n_nodes() == n
        assert R.L.refslam_solved_until(R.h) == F.solved_until() == n - 1 and R.L.refslam_n_kidnaps(R.h) == 1 and R.L.refslam_kidnap_status(R.h) == 0
        assert [R.L.refslam_world_setid(R.h, w) for w in range(2)] == [F.world_setid(w) for w in range(2)] == [0, 0]
        T = np.zeros((4, 4)); Tm = np.zeros((4, 4)); fq, ft = F.poses()
        for i in range(n):
            assert R.L.refslam_node_stamp(R.h, i) == int(g["stamps"][i]) and R.L.refslam_which_world(R.h, int(g["stamps"][i])) == F.which_world(int(g["stamps"][i])) == world[i]
            R.L.refslam_get_node_pose(R.h, i, T.ctypes.data_as(dp))          # the restored optimisation variable (ws_T_w * w_T_c again)
            assert np.allclose(T, pgo.pose_to_mat4(fq[i], ft[i]), rtol=0, atol=1e-9)
            R.L.refslam_manager_node_pose(R.h, i, Tm.ctypes.data_as(dp))     # the manager's pose: back in the keyframe's own world
            if world[i] == 0:
                assert np.allclose(Tm, T, rtol=0, atol=1e-9)
        # ---- the restored sessions go on: new keyframes in world 1 and a loop edge into the restored part
        last = int(g["stamps"][n - 1]); qn, tn = g["q"][n - 1], g["t"][n - 1]
        new_stamps = [last + (k + 1) * 10**8 for k in range(6)]
        new_q = np.array([pgo.quat_plus(qn, [0, 0, 0.01 * (k + 1)]) for k in range(6)]); new_t = np.array([tn + [0.5 * (k + 1), 0.1 * k, 0] for k in range(6)])
        R.add_nodes(new_stamps, new_q, new_t); F.add_nodes(new_stamps, new_q, new_t)
        all_stamps = np.r_[g["stamps"][:n], new_stamps]
        lq, lt = pgo.mat4_to_pose(pgo.inv4(pgo.pose_to_mat4(g["q"][n - 20], g["t"][n - 20])) @ pgo.pose_to_mat4(new_q[4], new_t[4]))
        R.add_loop_edges(all_stamps, [n + 4], [n - 20], [lq], [lt], [1.0]); F.add_loop_edges([n + 4], [n - 20], [lq], [lt], [1.0])
        assert R.wakeup() and F.solve_once()
        B = R.blocks(); od, lo, rg = B["type"] == 0, B["type"] == 1, B["type"] == 2
        o, l, r = F.alternative_terms(0), F.alternative_terms(1), F.reg_terms()
        assert np.array_equal(B["c1"][od], o["c1"]) and np.array_equal(B["c2"][od], o["c2"]) and np.allclose(B["w"][od], o["weight"], rtol=1e-9)
        assert B["c1"][od].min() == n and np.allclose(B["obs"][od], mats(o["obs_rot"], o["obs_t"]), rtol=0, atol=1e-9)   # odometry only from solvedUntil+1 on (:1570)
        assert np.array_equal(B["c1"][lo], l["c1"]) and np.array_equal(B["c2"][lo], l["c2"]) and len(l["c1"]) == 1
        assert np.array_equal(B["c1"][rg], r["node"]) and np.allclose(B["w"][rg], r["w"]) and np.allclose(B["obs"][rg], mats(r["q"], r["t"]), rtol=0, atol=1e-9)
        rq_, rt_, rs_, const = R.variables(); fq, ft = F.poses()
        assert np.allclose(rt_, ft, rtol=0, atol=1e-8) and same_quats(rq_, fq)
        assert const[:n].all() and not const[n:].any()                       # the restored keyframes are constant parameter blocks (:143-144), the new ones are free
    finally:
        R.close(); F.close()


def test_reference_node_with_a_solver_behind_its_ceres_solve_call_matches_the_python_front_end_session():
    """The drop-in, on the CPU: the reference's PoseGraphSLAM.cpp runs unmodified and every ceres::Solve it issues is served
    by a callback that reads the recorded problem, minimises it with the oracle's LM and writes the result into the
    reference's own optimisation arrays.  Over three wake-ups of a growing session the reference node must then hold the
    same poses and switches as the oracle's Python front-end driving the same LM — front-end rules, warm starts and
    re-anchored regularisers under REAL solves.  (tools/reference_node_with_libpgs.py is the same with libpgs.so on a GPU.)"""
    g = synth.generate_config(2, n_nodes=150, n_loop=24)
    order = np.argsort(np.maximum(g["la"], g["lb"]), kind="stable")
    R = ReferenceNode(); M = frontend.Manager(); P = frontend.ReferenceFrontEnd(M, odom_fanout=5)
    R.L.refslam_set_solve_callback.argtypes = [C.c_void_p, C.c_void_p]; R.L.refslam_write_vars.argtypes = [C.c_void_p, dp, dp, dp]
    solves = []

    def serve_ceres_solve():
        B = R.blocks(); q, t, s, _ = R.variables()
        od, lo, rg = B["type"] == 0, B["type"] == 1, B["type"] == 2
        pose = lambda Ms: (np.array([pgo.mat4_to_pose(X)[0] for X in Ms]).reshape(-1, 4), np.array([pgo.mat4_to_pose(X)[1] for X in Ms]).reshape(-1, 3))
        S = pgo.Problem(); S.set_nodes(q, t)
        oq, ot = pose(B["obs"][od]); S.add_odom_edges(B["c1"][od], B["c2"][od], oq, ot, B["w"][od])
        lq, lt = pose(B["obs"][lo]); S.add_loop_edges(B["c1"][lo], B["c2"][lo], lq, lt, B["w"][lo], s_init=s[B["sw"][lo]])
        rq, rt = pose(B["obs"][rg]); S.set_regularizers(B["c1"][rg], rq, rt, B["w"][rg])
        solves.append(S.solve())
        q2, t2 = S.poses(); s2 = s.copy(); s2[B["sw"][lo]] = S.switches()
        q2, t2, s2 = (np.ascontiguousarray(x, dtype=np.float64) for x in (q2, t2, s2))
        R.L.refslam_write_vars(R.h, q2.ctypes.data_as(dp), t2.ctypes.data_as(dp), s2.ctypes.data_as(dp))

    cb = C.CFUNCTYPE(None)(serve_ceres_solve)
    R.L.refslam_set_solve_callback(R.h, C.cast(cb, C.c_void_p))
    try:
        epos = 0; T = np.zeros((4, 4))
        for lo_ in range(0, 150, 50):
            sl = slice(lo_, lo_ + 50)
            R.add_nodes(g["stamps"][sl], g["q"][sl], g["t"][sl])
            for i in range(lo_, lo_ + 50):
                M.add_node(int(g["stamps"][i]), g["q"][i], g["t"][i])
            take = []
            while epos < len(order) and max(g["la"][order[epos]], g["lb"][order[epos]]) < lo_ + 50:
                take.append(order[epos]); epos += 1
            take = np.array(take, dtype=int)
            assert len(take)
            R.add_loop_edges(g["stamps"], g["la"][take], g["lb"][take], g["lq"][take], g["lt"][take], g["lw"][take])
            for e in take:
                M.add_loop_edge(int(g["la"][e]), int(g["lb"][e]), g["lq"][e], g["lt"][e], float(g["lw"][e]))
            assert R.wakeup()
            sp = P.trigger(solve=True)
            assert len(solves[-1]["iterations"]) == len(sp["iterations"]) and abs(solves[-1]["final_cost"] - sp["final_cost"]) <= 1e-9 * sp["final_cost"]
            for i in range(lo_ + 50):
                R.L.refslam_get_node_pose(R.h, i, T.ctypes.data_as(dp))
                assert np.allclose(T, pgo.pose_to_mat4(P.opt_q[i], P.opt_t[i]), rtol=0, atol=1e-8)
            assert R.L.refslam_solved_until(R.h) == P.solved_until == lo_ + 49
        assert len(solves) == 3 and solves[-1]["final_cost"] < solves[-1]["initial_cost"]
    finally:
        R.close()


@pytest.mark.parametrize("seed", range(24))
def test_random_sessions_match_the_reference_front_end(seed):
    """Differential test over random sessions (24 here; 60 were run): random chunk sizes, loop edges between random keyframes (inside a world,
    across the two worlds in both directions, and touching dead-zone keyframes, which every implementation must ignore),
    an optional kidnap at a random place, wake-ups after every chunk whether or not anything new arrived.  Even seeds keep
    the variables where the front end put them (the facade takes part); odd seeds move them after every wake-up with the
    stand-in solve (reference vs Python front-end)."""
    rng = np.random.default_rng(1000 + seed)
    moved = seed % 2 == 1
    g = synth.generate_config(4, n_nodes=int(rng.integers(30, 70)), n_interworld=4)
    stamps, k0, k1 = g["stamps"], g["k0"], g["k1"]
    two_worlds = rng.random() < 0.75
    last = int(np.nonzero(stamps <= (k0[1] if two_worlds else k0[0]))[0][-1])          # keyframes 0..last: world 0 [, dead zone, world 1]
    R = ReferenceNode(); M = frontend.Manager(); P = frontend.ReferenceFrontEnd(M, odom_fanout=5)
    F = None if moved else facade.Facade(odom_fanout=5, dry_run=True)
    if moved:
        R.L.refslam_set_perturb(R.h, 0.2)
    try:
        pos, kidnapped_sent, unkidnapped_sent, wake = 0, False, False, 0
        have = []                                                                      # keyframes ingested so far
        while pos <= last:
            hi = min(last + 1, pos + int(rng.integers(5, 40)))
            for i in range(pos, hi):
                if two_worlds and not kidnapped_sent and stamps[i] > k0[0]:
                    R.kidnap(int(k0[0]), 1); M.kidnap_indicator(int(k0[0]), 1); F and F.kidnap_indicator(int(k0[0]), 1); kidnapped_sent = True
                if two_worlds and kidnapped_sent and not unkidnapped_sent and stamps[i] > k1[0]:
                    R.kidnap(int(k1[0]), 0); M.kidnap_indicator(int(k1[0]), 0); F and F.kidnap_indicator(int(k1[0]), 0); unkidnapped_sent = True
                R.add_nodes(stamps[i:i + 1], g["q"][i:i + 1], g["t"][i:i + 1]); M.add_node(int(stamps[i]), g["q"][i], g["t"][i])
                F and F.add_nodes(stamps[i:i + 1], g["q"][i:i + 1], g["t"][i:i + 1])
                have.append(i)
            pos = hi
            for _ in range(int(rng.integers(0, 4))):
                a, b = (int(x) for x in rng.choice(have, size=2, replace=False)) if len(have) > 1 else (0, 0)
                if a == b:
                    continue
                wa, wb = M.which_world_is_this(int(stamps[a])), M.which_world_is_this(int(stamps[b]))
                if wa >= 0 and wb >= 0 and wa != wb and (wb, wa) != (0, 1) and not M.worlds.is_exist(wa, wb):
                    a, b = b, a                                                        # the first cross edge fixes key (0,1): later look-ups stay direct (no BFS branch)
                T = pgo.inv4(pgo.pose_to_mat4(g["q"][b], g["t"][b])) @ pgo.pose_to_mat4(g["q"][a], g["t"][a])
                q, t = pgo.mat4_to_pose(T); t = t + rng.normal(size=3) * 0.05
                R.add_loop_edges(stamps, [a], [b], [q], [t], [1.0]); M.add_loop_edge(a, b, q, t, 1.0); F and F.add_loop_edges([a], [b], [q], [t], [1.0])
            expect = P.prev_loopedge_len != len(M.edges) and not M.kidnapped          # the trigger condition (:1306-1319)
            fired = R.wakeup()
            P.trigger(solve=False)
            assert fired == expect
            if F is not None:
                assert F.solve_once() == fired
            if fired:
                wake += 1
                compare(R, F, P, 5)
                if moved:
                    stand_in_solve(P, 0.2, wake)
    finally:
        R.close(); F and F.close()



def test_loop_edge_time_stamp_lookup_matches_the_reference_callback():
    """loopclosure_pose_callback finds the two keyframes of a LoopEdge message by time stamp (src/NodeDataManager.cpp:107-189,
    find_indexof_node :274-299): stamps that are a little off still match, stamps further off drop the edge.  Random offsets
    around that tolerance, between keyframes 0.1 s apart, through the reference's callback and the product's."""
    rng = np.random.default_rng(21)
    n = 40
    stamps = 10**9 + np.arange(n, dtype=np.int64) * 10**8
    q = np.tile([0, 0, 0, 1.0], (n, 1)); t = np.c_[np.arange(n, dtype=float), np.zeros(n), np.zeros(n)]
    R = ReferenceNode(); F = facade.Facade(dry_run=True)
    R.L.refslam_edge.argtypes = [C.c_void_p, C.c_int, ip, ip, dp, dp]
    try:
        R.add_nodes(stamps, q, t); F.add_nodes(stamps, q, t)
        kept = 0
        for k in range(120):
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            off = [int(x) for x in rng.choice([0, 1, -1, 400_000, -400_000, 999_000, -999_000, 1_000_000, -1_000_000, 1_001_000, -1_001_000, 3_000_000, 49_000_000, -60_000_000], size=2)]
            sa, sb = int(stamps[a]) + off[0], int(stamps[b]) + off[1]
            lq = np.array([0, 0, 0, 1.0]); lt = np.array([float(k), 0.0, 0.0])
            before = R.L.refslam_n_edges(R.h)
            after = R.L.refslam_add_loop_edge(R.h, sa, sb, lq.ctypes.data_as(dp), lt.ctypes.data_as(dp), 1.0)
            got = F.add_loop_edge_stamped(sa, sb, lq, lt, 1.0)
            assert (after - before) == (1 if got else 0), (k, off, after - before, got)
            if got:
                ia, ib = C.c_int(), C.c_int(); T = np.zeros((4, 4)); w = C.c_double()
                R.L.refslam_edge(R.h, after - 1, C.byref(ia), C.byref(ib), T.ctypes.data_as(dp), C.byref(w))
                assert (ia.value, ib.value) == (a, b) and T[0, 3] == float(k)
                kept += 1
        assert 10 < kept < 110                                                 # both outcomes occurred
        assert F.n_loop == kept
    finally:
        R.close(); F.close()


@pytest.mark.parametrize("seed", range(3))
def test_world_indexing_with_several_kidnaps_matches_the_reference_manager(seed):
    """which_world_is_this / n_worlds / nodeidx_of_world_i_started / _ended (src/NodeDataManager.cpp:1127-1304) with up to four
    kidnaps, queried for stamps before, inside and after every world and dead zone, also while still kidnapped."""
    rng = np.random.default_rng(300 + seed)
    R = ReferenceNode(); F = facade.Facade(dry_run=True); M = frontend.Manager()
    try:
        t = 10**9; all_stamps = []; marks = []
        n_kid = int(rng.integers(2, 5)); end_kidnapped = seed == 2
        for w in range(n_kid + 1):
            for _ in range(int(rng.integers(3, 9))):                             # keyframes of world w
                t += int(rng.integers(5, 20)) * 10**7
                R.add_nodes([t], [[0, 0, 0, 1.0]], [[float(len(all_stamps)), 0, 0]]); F.add_nodes([t], [[0, 0, 0, 1.0]], [[float(len(all_stamps)), 0, 0]])
                M.add_node(t, np.array([0, 0, 0, 1.0]), np.array([float(len(all_stamps)), 0, 0])); all_stamps.append(t)
            if w == n_kid:
                break
            t += 3 * 10**7; k0 = t; marks.append(k0)
            R.kidnap(k0, 1); F.kidnap_indicator(k0, 1); M.kidnap_indicator(k0, 1)
            for _ in range(int(rng.integers(0, 4))):                             # keyframes that arrive while kidnapped
                t += int(rng.integers(5, 20)) * 10**7
                R.add_nodes([t], [[0, 0, 0, 1.0]], [[0.0, 1.0, 0]]); F.add_nodes([t], [[0, 0, 0, 1.0]], [[0.0, 1.0, 0]]); M.add_node(t, np.array([0, 0, 0, 1.0]), np.array([0, 1.0, 0])); all_stamps.append(t)
            if end_kidnapped and w == n_kid - 1:
                break
            t += 3 * 10**7; marks.append(t)
            R.kidnap(t, 0); F.kidnap_indicator(t, 0); M.kidnap_indicator(t, 0)
        probes = sorted(set(all_stamps + marks + [m + d for m in marks for d in (-1, 1, 10**6)] + [all_stamps[0] - 10**8, t + 10**9]
                            + [int(x) for x in rng.integers(all_stamps[0] - 10**8, t + 10**8, size=60)]))
        assert [R.L.refslam_which_world(R.h, s) for s in probes] == [F.which_world(s) for s in probes] == [M.which_world_is_this(s) for s in probes]
        nw = R.L.refslam_n_worlds(R.h)
        assert nw == F.n_worlds() == M.n_worlds()
        for w in range(-1, nw + 2):
            assert R.L.refslam_world_start(R.h, w) == F.world_start(w) == M.nodeidx_of_world_i_started(w), w
            assert R.L.refslam_world_end(R.h, w) == F.world_end(w) == M.nodeidx_of_world_i_ended(w), w
    finally:
        R.close(); F.close()


def test_pose_assembly_matches_the_reference_composer_thread():
    """Composer::pose_assember_thread (src/Composer.cpp:10-263), the reference's own code on its own thread, against the
    oracle's restatement (oracle/composer.py — which the device kernel of include/pgs_compose.h is held to in
    tests/test_composer.py): before any solve, after a solve that moved the keyframes of world 0, with keyframes that
    arrived after it, through a dead zone, in a second world before and after it is merged into the first."""
    from oracle import composer
    rng = np.random.default_rng(15)
    g = synth.generate_config(4, n_nodes=50, n_interworld=6)
    stamps, k0, k1 = g["stamps"], g["k0"], g["k1"]
    w0 = np.nonzero(stamps <= k0[0])[0]; dead = np.nonzero((stamps > k0[0]) & (stamps <= k1[0]))[0]; w1 = np.nonzero((stamps > k1[0]) & (stamps <= k0[1]))[0]
    R = ReferenceNode(); M = frontend.Manager(); P = frontend.ReferenceFrontEnd(M, odom_fanout=5)
    R.L.refslam_compose_once.argtypes = [C.c_void_p, dp, C.c_int]
    R.L.refslam_set_perturb(R.h, 0.3)
    checked = [0]

    def nodes(idx):
        R.add_nodes(stamps[idx], g["q"][idx], g["t"][idx])
        for i in idx:
            M.add_node(int(stamps[i]), g["q"][i], g["t"][i])

    def loops(pairs):
        for a, b in pairs:
            T = pgo.inv4(pgo.pose_to_mat4(g["q"][b], g["t"][b])) @ pgo.pose_to_mat4(g["q"][a], g["t"][a])
            q, t = pgo.mat4_to_pose(T); t = t + rng.normal(size=3) * 0.05
            R.add_loop_edges(stamps, [a], [b], [q], [t], [1.0]); M.add_loop_edge(a, b, q, t, 1.0)

    def wake(k):
        assert R.wakeup(); P.trigger(solve=False); stand_in_solve(P, 0.3, k)

    def check():
        n = len(M.poses)
        buf = np.zeros((n + 8, 4, 4))
        got = R.L.refslam_compose_once(R.h, buf.ctypes.data_as(dp), n + 8)
        slam = [pgo.pose_to_mat4(P.opt_q[i], P.opt_t[i]) for i in range(len(P.opt_q))]
        want, wid, _ = composer.assemble(M, slam, P.solved_until)
        assert got == n == len(want)
        assert np.allclose(buf[:n], want, rtol=0, atol=1e-8), int(np.abs(buf[:n] - want).reshape(n, -1).max(axis=1).argmax())
        checked[0] += 1

    try:
        nodes(w0[:30]); check()                                               # nothing solved: odometry
        loops([(int(w0[25]), int(w0[3]))]); wake(1); check()                  # world 0 solved and moved
        nodes(w0[30:]); check()                                               # keyframes after the solve: carried forward
        R.kidnap(int(k0[0]), 1); M.kidnap_indicator(int(k0[0]), 1); nodes(dead); check()          # dead zone hangs off the last pose of world 0
        R.kidnap(int(k1[0]), 0); M.kidnap_indicator(int(k1[0]), 0); nodes(w1[:25]); check()       # a new world, not solved, not connected
        loops([(int(w1[20]), int(w1[2]))]); wake(2); check()                  # world 1 solved on its own
        nodes(w1[25:]); loops([(int(w1[30]), int(w0[10]))]); wake(3); check() # merged into world 0's set
        nodes(np.arange(int(w1[-1]) + 1, int(w1[-1]) + 1)); check()
        assert checked[0] == 8
    finally:
        R.close()


def test_solved_posegraph_json_written_and_restored_by_the_reference_composer(tmp_path):
    """Composer::saveStateToDisk of the reference writes solved_posegraph.json for a two-world session (its own assembler
    filled global_lmb); the product writes the same session.  KidnapTimestamps and WorldsData must agree, the reference's
    SolvedPoseGraph must be readable by the product's reader, and — the other way round — Composer::loadStateFromDisk of the
    reference itself (not a restatement of its call sequence) must restore a session from the file the PRODUCT restores from."""
    import json
    rng = np.random.default_rng(16)
    R = ReferenceNode(); F = facade.Facade(odom_fanout=5, dry_run=True)
    for f, a in dict(refslam_compose_once=[dp, C.c_int], refslam_composer_save=[C.c_char_p], refslam_composer_load=[C.c_char_p], refslam_slam_n_nodes=[],
                     refslam_n_kidnaps=[], refslam_kidnap_status=[]).items():
        getattr(R.L, f).argtypes = [C.c_void_p] + a
    dR, dF = tmp_path / "ref", tmp_path / "ours"; dR.mkdir(); dF.mkdir()
    try:
        g, n = _two_world_session(R, F, rng)
        lmb = np.zeros((n, 4, 4))
        assert R.L.refslam_compose_once(R.h, lmb.ctypes.data_as(dp), n) == n
        assert R.L.refslam_composer_save(R.h, str(dR).encode()) == 1
        F.save_state_to_disk(dF)                                             # both end the current world at the last keyframe first (Composer.cpp:967-974)
        assert R.L.refslam_kidnap_status(R.h) == 1 and R.L.refslam_n_kidnaps(R.h) == 1
        Jr = json.load(open(dR / "solved_posegraph.json")); Jo = json.load(open(dF / "solved_posegraph.json"))
        assert sorted(Jr) == sorted(Jo) == ["KidnapTimestamps", "SolvedPoseGraph", "WorldsData"]
        assert Jr["KidnapTimestamps"] == Jo["KidnapTimestamps"] and len(Jr["KidnapTimestamps"]["kidnap_starts"]) == 2 and len(Jr["KidnapTimestamps"]["kidnap_ends"]) == 1
        assert Jr["WorldsData"]["disjoint_set"] == Jo["WorldsData"]["disjoint_set"] and Jr["WorldsData"]["vec_world_starts"] == Jo["WorldsData"]["vec_world_starts"]
        assert Jr["WorldsData"]["vec_world_ends"] == Jo["WorldsData"]["vec_world_ends"]
        assert [sorted(x) for x in Jr["WorldsData"]["rel_pose_between_worlds__wb_T_wa"]] == [sorted(x) for x in Jo["WorldsData"]["rel_pose_between_worlds__wb_T_wa"]]
        # the reference-written SolvedPoseGraph through the product's reader
        T, st, wid, sid = facade.io_load_solved_posegraph(dR / "solved_posegraph.json")
        assert len(T) == n and np.allclose(T, lmb, rtol=1e-15, atol=1e-300)
        assert list(st) == [int(s) for s in g["stamps"][:n]] and list(wid) == [F.which_world(int(s)) for s in g["stamps"][:n]]
        assert [int(x) for x in sid] == [F.world_setid(int(w)) if w >= 0 else -1 for w in wid]
        entry = Jr["SolvedPoseGraph"][3]
        assert sorted(entry) == ["seq", "setID_of_worldID", "stampNSec", "w_T_c", "worldID"] and sorted(entry["w_T_c"]) == ["cols", "data", "data_pretty", "rows"]
        assert entry["w_T_c"]["data_pretty"] == facade.io_prettyprint(lmb[3]) and entry["w_T_c"]["data"] == facade.io_mat_to_string(lmb[3], solved_layout=True)
    finally:
        R.close(); F.close()
    # ---- Composer::loadStateFromDisk itself, on the reference-written file and on the same file with the product's WorldsData / KidnapTimestamps
    Jmix = dict(Jr); Jmix["WorldsData"] = Jo["WorldsData"]; Jmix["KidnapTimestamps"] = Jo["KidnapTimestamps"]
    dM = tmp_path / "mixed"; dM.mkdir(); json.dump(Jmix, open(dM / "solved_posegraph.json", "w"), indent=4)
    for d in (dR, dM):
        R2 = ReferenceNode(); F2 = facade.Facade(odom_fanout=5, dry_run=True)
        for f, a in dict(refslam_composer_load=[C.c_char_p], refslam_slam_n_nodes=[], refslam_n_kidnaps=[], refslam_kidnap_status=[]).items():
            getattr(R2.L, f).argtypes = [C.c_void_p] + a
        try:
            assert R2.L.refslam_composer_load(R2.h, str(d).encode()) == 1
            F2.load_state_from_disk(d)
            assert R2.L.refslam_n_nodes(R2.h) == F2.n_keyframes() == n and R2.L.refslam_slam_n_nodes(R2.h) == F2.n_nodes() == n
            assert R2.L.refslam_solved_until(R2.h) == F2.solved_until() == n - 1 and R2.L.refslam_n_kidnaps(R2.h) == 1 and R2.L.refslam_kidnap_status(R2.h) == 1
            fq, ft = F2.poses(); T = np.zeros((4, 4))
            for i in range(n):
                R2.L.refslam_get_node_pose(R2.h, i, T.ctypes.data_as(dp))
                assert np.allclose(T, pgo.pose_to_mat4(fq[i], ft[i]), rtol=0, atol=1e-9)
            # the restored session receives its first keyframe: the kidnap that the save started ends there and a NEW world begins
            # (NodeDataManager.cpp:74-92); a loop edge from it into restored world 0 merges that world into the old set
            last = int(g["stamps"][n - 1])
            new_stamps = [last + (k + 1) * 10**8 for k in range(8)]
            new_q = np.tile([0, 0, 0, 1.0], (8, 1)); new_t = np.c_[np.arange(8) * 0.7, np.zeros(8), np.zeros(8)]
            R2.add_nodes(new_stamps, new_q, new_t); F2.add_nodes(new_stamps, new_q, new_t)
            assert R2.L.refslam_kidnap_status(R2.h) == 0 and R2.L.refslam_n_kidnaps(R2.h) == 2 and R2.L.refslam_n_worlds(R2.h) == F2.n_worlds() == 3
            wr_, wf_ = [R2.L.refslam_which_world(R2.h, int(s_)) for s_ in new_stamps], [F2.which_world(int(s_)) for s_ in new_stamps]
            assert wr_ == wf_, (wr_, wf_)
            assert wr_ == [-2] + [2] * 7        # the keyframe AT the un-kidnap stamp still counts as dead zone (<=, NodeDataManager.cpp:1127-1198)
            all_stamps = np.r_[g["stamps"][:n], new_stamps]
            lq, lt = np.array([0, 0, 0, 1.0]), np.array([0.3, -0.2, 0.1])
            R2.add_loop_edges(all_stamps, [n + 5], [7], [lq], [lt], [1.0]); F2.add_loop_edges([n + 5], [7], [lq], [lt], [1.0])
            assert R2.wakeup() and F2.solve_once()
            B = R2.blocks(); od, lo, rg = B["type"] == 0, B["type"] == 1, B["type"] == 2
            o, l, r = F2.alternative_terms(0), F2.alternative_terms(1), F2.reg_terms()
            assert np.array_equal(B["c1"][od], o["c1"]) and np.array_equal(B["c2"][od], o["c2"]) and np.allclose(B["w"][od], o["weight"], rtol=1e-9)
            assert np.array_equal(B["c1"][lo], l["c1"]) and np.array_equal(B["c2"][lo], l["c2"]) and np.allclose(B["obs"][lo], mats(l["obs_rot"], l["obs_t"]), rtol=0, atol=1e-9)
            assert np.array_equal(B["c1"][rg], r["node"]) and np.allclose(B["w"][rg], r["w"]) and np.allclose(B["obs"][rg], mats(r["q"], r["t"]), rtol=0, atol=1e-9)
            rq_, rt_, rs_, const = R2.variables(); fq, ft = F2.poses()
            assert np.allclose(rt_, ft, rtol=0, atol=1e-8) and same_quats(rq_, fq) and const[:n].all() and not const[n:].any()
            assert [R2.L.refslam_world_setid(R2.h, w) for w in range(3)] == [F2.world_setid(w) for w in range(3)]
        finally:
            R2.close(); F2.close()
