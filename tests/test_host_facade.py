"""CPU tests (no GPU): the ROS-free facade's graph-construction rules against the oracle's Python
front-end restatement of reference src/PoseGraphSLAM.cpp:1287-1950, world bookkeeping samples from
the reference's scratch mains (src/test_disjointset.cpp:26-44, src/test_bfs.cpp:95-121) turned into
assertions, and the C-ABI export check."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import solve_keyframe_pose_graph_b200 as pgs
from oracle import frontend, pgo
from solve_keyframe_pose_graph_b200 import facade, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = pgs.lib()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
        names = sorted(set(re.findall(r"\b(pgs_[a-z_0-9]+)\s*\(", txt)))
        assert names, hdr
        for n in names:
            assert hasattr(L, n), f"{n} declared in {os.path.basename(hdr)} is not exported by libpgs.so"


def test_create_fails_loudly_without_a_device_or_reports_ok():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pgs.PgsError, match="no usable CUDA device|CUDA"):
        pgs.PoseGraphSolver()


def test_disjoint_set_rank_rule():
    # merging (3,2) then (2,0) makes 2 the root (rank 1 beats rank 0)
    d = frontend.DisjointSetForest()
    for i in range(4):
        d.add_element(i)
    d.union_sets(3, 2)
    assert d.find_set(3) == 2 and d.find_set(2) == 2
    d.union_sets(2, 0)
    assert d.find_set(0) == 2 and d.set_count() == 2
    # facade agrees: worlds 0..3, edges (3,2) then (2,0)
    F = facade.Facade(dry_run=True)
    I = np.array([0, 0, 0, 1.0]); z = np.zeros(3)
    stamps = np.arange(8, dtype=np.int64) * 10**8 + 10**9
    for w in range(4):
        F.add_nodes(stamps[2 * w:2 * w + 2], np.tile(I, (2, 1)), np.zeros((2, 3)))
        if w < 3:
            F.kidnap_indicator(stamps[2 * w + 1], 1); F.kidnap_indicator(stamps[2 * w + 1] + 10**7, 0)
    assert F.n_worlds() == 4
    assert [F.which_world(s) for s in stamps] == [0, 0, 1, 1, 2, 2, 3, 3]
    F.add_loop_edges([6], [4], [I], [z], [1.0])   # world 3 -> world 2
    assert F.solve_once()
    assert [F.world_setid(w) for w in range(4)] == [0, 1, 2, 2]
    F.add_loop_edges([4], [0], [I], [z], [1.0])   # world 2 -> world 0
    assert F.solve_once()
    assert [F.world_setid(w) for w in range(4)] == [2, 1, 2, 2]


def test_bfs_sample_from_reference_scratch_main():
    # src/test_bfs.cpp:95-121: edges 0->1,0->2,1->2,2->0,2->3,4->5,4->6,5->6; BFS(2); path from 0
    edges = [(0, 1), (0, 2), (1, 2), (2, 0), (2, 3), (4, 5), (4, 6), (5, 6)]
    parent, visited = frontend.bfs_parents(7, edges, 2)
    assert parent[0] == 2 and parent[3] == 2 and parent[1] == 0 and parent[2] == -2
    assert visited[:4] == [True] * 4 and visited[4:] == [False] * 3
    assert frontend.path_from(parent, visited, 0) == [0, 2]
    assert frontend.path_from(parent, visited, 5) == []


def test_pose_between_worlds_chained_by_bfs():
    rng = np.random.default_rng(3)
    def rp():
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        return pgo.pose_to_mat4(q, rng.normal(size=3) * 10)
    W = frontend.Worlds()
    for i in range(4):
        W.world_starts(i)
    T10, T21, T32 = rp(), rp(), rp()
    W.setPoseBetweenWorlds(1, 0, T10); W.setPoseBetweenWorlds(2, 1, T21); W.setPoseBetweenWorlds(3, 2, T32)
    assert np.allclose(W.getPoseBetweenWorlds(3, 0), T32 @ T21 @ T10, atol=1e-12)
    assert np.allclose(W.getPoseBetweenWorlds(0, 2), np.linalg.inv(T21 @ T10), atol=1e-9)
    assert (3, 0) in W.rel        # memoised


@pytest.mark.parametrize("config,kw,fan", [(1, {}, 1), (2, dict(n_nodes=600, n_loop=80), 3), (2, dict(n_nodes=300, n_loop=20), 5)])
def test_first_trigger_matches_oracle_frontend(config, kw, fan):
    g = synth.generate_config(config, **kw)
    F = facade.Facade(odom_fanout=fan, dry_run=True); F.ingest(g)
    M = frontend.Manager(); M.ingest(g)
    R = frontend.ReferenceFrontEnd(M, odom_fanout=fan)
    assert F.solve_once() and R.trigger(solve=False) is None
    o = F.odom_terms()
    assert len(o["u"]) == len(R.odom) == sum(min(fan, u) for u in range(g["N"]))
    assert np.array_equal(o["u"], [x[0] for x in R.odom]) and np.array_equal(o["umf"], [x[1] for x in R.odom])
    assert np.allclose(o["w"], [x[4] for x in R.odom], rtol=1e-9, atol=0)
    assert np.allclose(o["t"], np.array([x[3] for x in R.odom]), atol=1e-9)
    dq = np.abs(np.sum(o["q"] * np.array([x[2] for x in R.odom]), axis=1))
    assert np.all(dq > 1 - 1e-12)
    r = F.reg_terms()
    assert list(r["node"]) == [x[0] for x in R.regs] == [0]
    assert np.allclose(r["w"], [x[3] for x in R.regs]) and np.isclose(r["w"][0], max(1.1, np.log(g["N"]) / 2))
    q, t = F.poses()
    assert np.allclose(t, np.array(R.opt_t), atol=1e-9) and np.allclose(t, g["t"], atol=1e-9)   # first trigger: guesses = odometry poses
    assert F.solved_until() == g["N"] - 1 and F.status() == 3
    assert not F.solve_once()         # no new loop edge -> no trigger (PoseGraphSLAM.cpp:1306-1312)


def test_multi_world_trigger_matches_oracle_frontend():
    # config-4 recipe at reduced size: 4 worlds x 150 nodes, 5 dead-zone nodes between, 24 inter-world edges
    g = synth.generate_config(4, n_nodes=150, n_interworld=24)
    assert g["N"] == 4 * 150 + 3 * 5 and len(g["k0"]) == 3
    F = facade.Facade(odom_fanout=3, dry_run=True); F.ingest(g)
    M = frontend.Manager(); M.ingest(g)
    R = frontend.ReferenceFrontEnd(M, odom_fanout=3)
    assert [F.which_world(s) for s in g["stamps"]] == [M.which_world_is_this(int(s)) for s in g["stamps"]]
    ww = np.array([M.which_world_is_this(int(s)) for s in g["stamps"]])
    assert (ww < 0).sum() == 15 and set(ww[ww >= 0]) == {0, 1, 2, 3}
    assert F.solve_once(); R.trigger(solve=False)
    assert [F.world_setid(w) for w in range(4)] == [M.worlds.find_setID_of_world_i(w) for w in range(4)]
    assert [F.world_start(w) for w in range(4)] == [M.nodeidx_of_world_i_started(w) for w in range(4)] == [0, 155, 310, 465]
    assert [F.world_end(w) for w in range(4)] == [M.nodeidx_of_world_i_ended(w) for w in range(4)] == [149, 304, 459, 614]
    o = F.odom_terms()
    assert np.array_equal(o["u"], [x[0] for x in R.odom]) and np.array_equal(o["umf"], [x[1] for x in R.odom])
    # 5 dead-zone nodes >= fan-out 3: no odometry edge may cross worlds
    assert np.all(ww[o["u"]] == ww[o["umf"]]) and np.all(ww[o["u"]] >= 0)
    r = F.reg_terms()
    assert list(r["node"]) == [x[0] for x in R.regs] and np.allclose(r["w"], [x[3] for x in R.regs])
    q, t = F.poses()
    assert np.allclose(t, np.array(R.opt_t), atol=1e-7)
    for m_ in range(4):
        for n_ in range(4):
            A = F.pose_between_worlds(m_, n_)
            assert A is not None and np.allclose(A, M.worlds.getPoseBetweenWorlds(m_, n_), atol=1e-7)
    # every world was mapped into the set root's frame: the first inter-world edges are inliers, so the
    # initial guess of connected nodes must be consistent with the loop observation up to odometry drift
    root = F.world_setid(0)
    assert all(F.world_setid(w) == root for w in range(4))


def test_stamped_loop_edge_lookup_and_drop():
    F = facade.Facade(dry_run=True)
    I = np.array([0, 0, 0, 1.0])
    stamps = np.arange(10, dtype=np.int64) * 10**8 + 10**9
    F.add_nodes(stamps, np.tile(I, (10, 1)), np.zeros((10, 3)))
    assert F.add_loop_edge_stamped(stamps[7] + 900_000, stamps[2] - 900_000, I, np.zeros(3))      # within 1 ms
    assert not F.add_loop_edge_stamped(stamps[7] + 1_000_000, stamps[2], I, np.zeros(3))           # exactly 1 ms: no match
    assert not F.add_loop_edge_stamped(stamps[9] + 10**9, stamps[2], I, np.zeros(3))               # unknown keyframe: dropped


def test_generator_is_deterministic_and_matches_baseline_counts():
    for cfg, (n, el) in {1: (50, 1), 2: (10000, 2000)}.items():
        a = synth.generate_config(cfg); b = synth.generate_config(cfg)
        assert a["N"] == n and len(a["la"]) == el
        for k in ("q", "t", "lq", "lt", "la", "lb"):
            assert np.array_equal(a[k], b[k])
    s3 = synth.config_spec(3); s5 = synth.config_spec(5); s4 = synth.config_spec(4)
    assert (s3.n_nodes, s3.n_loop, s3.outlier_fraction) == (100000, 50000, 0.10)
    assert (s5.n_nodes, s5.n_loop) == (1000000, 500000)
    assert (s4.n_nodes, s4.n_worlds, s4.n_interworld, s4.deadzone_nodes) == (25000, 4, 200, 5)
    g = synth.generate_config(2)
    gap = g["la"] - g["lb"]
    assert gap.min() >= 50 and gap.max() <= min(10000 // 8, 2000)
    assert abs(np.linalg.norm(g["q"], axis=1) - 1).max() < 1e-12
    # odometry drift is small per step: consecutive relative translation ~ 1 m
    step = np.linalg.norm(np.diff(g["t"], axis=0), axis=1)
    assert abs(step.mean() - 1.0) < 0.01


def test_ros_message_callbacks_build_the_same_session_as_the_plain_ingest():
    # nav_msgs/Odometry, LoopEdge.msg and the kidnap Header routed through the reference's callback names
    g = synth.generate_config(4, n_nodes=40, n_worlds=3, n_interworld=8)
    A = facade.Facade(odom_fanout=3, dry_run=True); A.ingest(g)
    B = facade.Facade(odom_fanout=3, dry_run=True)
    ev = sorted([(int(s), "kidnapped") for s in g["k0"]] + [(int(s), "unkidnapped") for s in g["k1"]])
    k = 0
    for i in range(g["N"]):
        while k < len(ev) and ev[k][0] < g["stamps"][i]:
            assert B.rcvd_kidnap_indicator_callback(*ev[k]); k += 1
        B.camera_pose_callback(g["stamps"][i], g["t"][i], g["q"][i])
    assert not B.rcvd_kidnap_indicator_callback(g["stamps"][-1] + 5, "lost")          # the reference exits on anything else (:789-791)
    for e in range(len(g["la"])):                                                       # timestamp0 -> a, timestamp1 -> b, pose_1T0 = b_T_a
        assert B.loopclosure_pose_callback(g["stamps"][g["la"][e]], g["stamps"][g["lb"][e]], g["lt"][e], g["lq"][e], float(g["lw"][e]), "t")
    assert not B.loopclosure_pose_callback(g["stamps"][3] + 5 * 10**6, g["stamps"][1], g["lt"][0], g["lq"][0])   # no keyframe within 1 ms: dropped
    assert [A.which_world(s) for s in g["stamps"]] == [B.which_world(s) for s in g["stamps"]]
    assert A.solve_once() and B.solve_once()
    a, b = A.odom_terms(), B.odom_terms()
    assert np.array_equal(a["u"], b["u"]) and np.allclose(a["t"], b["t"], atol=1e-12) and np.allclose(a["w"], b["w"], rtol=1e-6)
    (qa, ta), (qb, tb) = A.poses(), B.poses()
    assert np.allclose(ta, tb, atol=1e-9) and [A.world_setid(w) for w in range(3)] == [B.world_setid(w) for w in range(3)]
    A.close(); B.close()


def test_load_state_restores_a_constant_backbone(tmp_path):
    # PoseGraphSLAM::load_state (src/PoseGraphSLAM.cpp:40-170) after a save / load round trip, dry run: variables of the
    # restored keyframes sit at ws_T_w * w_T_c, solvedUntil is the last of them, and the next trigger only adds odometry
    # edges for the keyframes that arrive afterwards — facade vs the oracle front-end
    g = synth.generate_config(4, n_nodes=50, n_worlds=3, n_interworld=9)
    A = facade.Facade(odom_fanout=3, dry_run=True); A.ingest(g); assert A.solve_once()          # merges the three worlds
    assert A.save_json(tmp_path) & 1
    M = frontend.Manager(); M.ingest(g)
    R0 = frontend.ReferenceFrontEnd(M, odom_fanout=3); R0.trigger(solve=False)                  # same merges in the oracle's Worlds
    F = facade.Facade(odom_fanout=3, dry_run=True)
    F.load_worlds_state(tmp_path / "solved_posegraph.json")                                     # Worlds first (Composer.cpp:1137), then the keyframes
    with pytest.raises(pgs.PgsError):
        F.load_state()                                                                           # no keyframes yet (the reference exits)
    F.close()
    F = facade.Facade(odom_fanout=3, dry_run=True); F.load_posegraph_json(tmp_path); F.n_loop = len(g["la"])
    R = frontend.ReferenceFrontEnd(M, odom_fanout=3)
    # worlds of F are not merged yet (log_posegraph.json carries no relative poses): nodes of worlds 1, 2 have set id == world id
    F.load_state(); R2 = frontend.ReferenceFrontEnd(frontend.Manager(), odom_fanout=3)
    assert F.solved_until() == g["N"] - 1 and F.n_nodes() == g["N"]
    q, t = F.poses()
    assert np.allclose(t, g["t"], atol=1e-12)                                                    # ws_T_w = I while the worlds are separate
    R.load_state()                                                                               # merged worlds: poses move into the set root's frame
    ww = np.array([M.which_world_is_this(int(s)) for s in g["stamps"]])
    moved = np.abs(np.array(R.opt_t) - g["t"]).max(axis=1) > 1e-9
    root = M.worlds.find_setID_of_world_i(0)
    assert R.solved_until == g["N"] - 1 and R.n_constant == g["N"] and not moved[ww == root].any() and moved[(ww >= 0) & (ww != root)].all()
    F.close(); A.close()


def test_solver_thread_with_concurrent_ingest_and_getters():
    # the reference's threading model (keyframe_pose_graph_slam_node.cpp:353,475-477): the solver polls on its own thread
    # while the ROS callbacks append keyframes / loop edges and the Composer reads poses.  Dry run, 500 Hz polling.
    import threading, time
    g = synth.generate_config(2, n_nodes=400, n_loop=60)
    order = np.argsort(np.maximum(g["la"], g["lb"]), kind="stable")
    F = facade.Facade(odom_fanout=3, dry_run=True)
    F.thread_start(500.0)
    stop = False; seen = []
    def reader():                                   # what the Composer / Viz threads do at 30 Hz
        while not stop:
            n = F.n_nodes()
            if n:
                q, t = F.poses(); assert len(t) >= n and np.isfinite(t).all()
            seen.append((F.status(), F.solved_until()))
    th = threading.Thread(target=reader); th.start()
    epos = 0
    for lo in range(0, 400, 50):
        F.add_nodes(g["stamps"][lo:lo + 50], g["q"][lo:lo + 50], g["t"][lo:lo + 50])
        take = []
        while epos < len(order) and max(g["la"][order[epos]], g["lb"][order[epos]]) < lo + 50:
            take.append(order[epos]); epos += 1
        if take:
            take = np.array(take); F.add_loop_edges(g["la"][take], g["lb"][take], g["lq"][take], g["lt"][take], g["lw"][take])
        time.sleep(0.02)
    time.sleep(0.1)
    for call in (F.odom_terms, F.reg_terms, lambda: F.alternative_terms(0)):      # the block lists belong to the running solver thread
        with pytest.raises(pgs.PgsError, match="while the solver thread runs"):
            call()
    n_solves = F.thread_stop(); stop = True; th.join()
    assert n_solves >= 2 and epos == len(order)
    assert F.solved_until() == 399 and F.n_nodes() == 400 and F.status() in (0, 3)
    assert all(s in (-1, 0, 1, 2, 3) for s, _ in seen) and [u for _, u in seen] == sorted(u for _, u in seen)   # solvedUntil never goes back
    # same graph as a single sequential trigger: same odometry terms (order aside) and the same final guesses for the last chunk
    G = facade.Facade(odom_fanout=3, dry_run=True); G.ingest(g); assert G.solve_once()
    a, b = F.odom_terms(), G.odom_terms()
    assert sorted(zip(a["u"].tolist(), a["umf"].tolist())) == sorted(zip(b["u"].tolist(), b["umf"].tolist()))
    F.close(); G.close()


def test_explicit_graph_api_add_odometry_edge_and_add_loop_edge():
    # north_star's explicit-graph entry points (absent in the reference): with derive_odometry off the problem contains
    # exactly the blocks that were added, in order, and a trigger fires on explicit odometry edges alone
    g = synth.generate_config(1)                                  # 50-node chain, one loop edge
    F = facade.Facade(odom_fanout=5, derive_odometry=False, dry_run=True)
    F.add_nodes(g["stamps"], g["q"], g["t"])
    assert not F.solve_once()                                     # nothing to solve yet
    rng = np.random.default_rng(0)
    added = []
    for u in range(1, g["N"]):
        q = rng.normal(size=4); q /= np.linalg.norm(q); t = rng.normal(size=3); w = float(rng.uniform(0.1, 1.0))
        F.add_odometry_edge(u, u - 1, q, t, w); added.append((u, u - 1, q, t, w))
    with pytest.raises(pgs.PgsError):
        F.add_odometry_edge(3, 99, [0, 0, 0, 1.0], [0, 0, 0.0], 1.0)   # node out of range
    assert F.solve_once()                                         # explicit odometry edges alone trigger
    o = F.odom_terms()
    assert list(o["u"]) == [a[0] for a in added] and list(o["umf"]) == [a[1] for a in added]
    assert np.allclose(o["w"], [a[4] for a in added]) and np.allclose(o["t"], np.array([a[3] for a in added]), atol=1e-12)
    assert np.all(np.abs(np.sum(o["q"] * np.array([a[2] for a in added]), axis=1)) > 1 - 1e-12)
    F.add_loop_edges(g["la"], g["lb"], g["lq"], g["lt"], g["lw"])  # addLoopEdge: stored in the manager, bound as (b, a, switch)
    assert F.solve_once() and len(F.odom_terms()["u"]) == len(added)   # no derived odometry was added behind our back
    F.close()


def test_c_abi_headers_compile_as_plain_c(tmp_path):
    # the drop-in boundary is a C ABI: every header under include/ must be valid C99 on its own (no C++, no torch types)
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text("".join(f'#include "{os.path.basename(h)}"\n' for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))) + "int main(void) { return 0; }\n")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "hdr.o")])


def test_loop_edge_onto_the_same_keyframe_builds_no_block_and_later_triggers_go_on():
    """both stamps of a loop message can resolve to one keyframe (1 ms lookup).  The manager keeps the
    edge, as the reference's does (src/NodeDataManager.cpp:107-189), but no residual block is built for it — Ceres
    itself refuses a block that binds one parameter block twice — and the triggers after it work as usual; the block
    lists never hold a term twice, however often the trigger runs."""
    F = facade.Facade(dry_run=True, odom_fanout=2)
    I = np.array([0, 0, 0, 1.0]); z = np.zeros(3)
    st = np.arange(10, dtype=np.int64) * 10**8 + 10**9
    F.add_nodes(st, np.tile(I, (10, 1)), np.cumsum(np.ones((10, 3)), axis=0))
    F.add_loop_edges([5, 7], [5, 2], [I, I], [z, z], [1.0, 1.0])
    assert F.solve_once()
    n_odom = len(F.odom_terms()["u"])
    assert n_odom == 2 * 10 - 3 and len(F.switches()) == 2
    F.add_loop_edges([9], [1], [I], [z], [1.0])
    assert F.solve_once() and F.status() == 3
    assert len(F.odom_terms()["u"]) == n_odom            # nothing new to derive, nothing derived twice
    assert not F.solve_once() and F.solve_once(force=True) and len(F.odom_terms()["u"]) == n_odom
    F.close()


@pytest.mark.parametrize("config,kw", [(2, dict(n_nodes=9000, n_loop=900)), (4, dict(n_nodes=2500, n_interworld=40))])
def test_trigger_on_several_host_threads_equals_the_sequential_trigger(config, kw, monkeypatch):
    """solve_once runs its per-keyframe rules (odometry terms, initial guesses) in contiguous chunks on the host's cores
    (csrc/host/PoseGraphSLAM.cpp for_chunks); the lists and guesses must be those of one thread, bit for bit — across
    worlds (config 4), and for a second wake-up that dead-reckons the new keyframes from the last solved one (config 2)."""
    g = synth.generate_config(config, **kw)
    runs = {}
    for threads in ("1", "5"):
        monkeypatch.setenv("PGS_HOST_THREADS", threads)
        F = facade.Facade(odom_fanout=3, dry_run=True)
        if config == 4:
            F.ingest(g)
            assert F.solve_once()
            snap = [F.odom_terms(), F.poses(), F.reg_terms()]
        else:
            cut = int(0.6 * g["N"])
            keep = (g["la"] < cut) & (g["lb"] < cut)
            F.add_nodes(g["stamps"][:cut], g["q"][:cut], g["t"][:cut])
            F.add_loop_edges(g["la"][keep], g["lb"][keep], g["lq"][keep], g["lt"][keep], g["lw"][keep])
            assert F.solve_once()
            snap = [F.odom_terms(), F.poses(), F.reg_terms()]
            F.add_nodes(g["stamps"][cut:], g["q"][cut:], g["t"][cut:])
            F.add_loop_edges(g["la"][~keep], g["lb"][~keep], g["lq"][~keep], g["lt"][~keep], g["lw"][~keep])
            assert F.solve_once()
            snap += [F.odom_terms(), F.poses(), F.reg_terms()]
        runs[threads] = snap
        F.close()
    for a, b in zip(runs["1"], runs["5"]):
        if isinstance(a, dict):
            assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
        else:
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert len(runs["1"][0]["u"]) > 3 * 2048   # several chunks were in play


def test_library_asks_for_more_hardware_queues_unless_the_host_chose():
    """The streams of a factorisation must not share a hardware work queue with each other ("Hardware work
    queues"): libpgs.so puts CUDA_DEVICE_MAX_CONNECTIONS=32 into the environment when it is loaded — before a CUDA context
    can exist in a host that links it — and leaves a value the host chose alone; the Python package does the same on import."""
    import subprocess, sys, os
    from solve_keyframe_pose_graph_b200 import capi
    getenv = ("import ctypes; L = ctypes.CDLL(None); L.getenv.restype = ctypes.c_char_p; ctypes.CDLL(%r); "
              "print(L.getenv(b'CUDA_DEVICE_MAX_CONNECTIONS').decode())") % capi.library_path()
    env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}
    assert subprocess.check_output([sys.executable, "-c", getenv], env=env).decode().strip() == "32"
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "4"
    assert subprocess.check_output([sys.executable, "-c", getenv], env=env).decode().strip() == "4"
    del env["CUDA_DEVICE_MAX_CONNECTIONS"]
    pkg = "import os, sys; sys.path.insert(0, %r); import solve_keyframe_pose_graph_b200; print(os.environ['CUDA_DEVICE_MAX_CONNECTIONS'])" % ROOT
    assert subprocess.check_output([sys.executable, "-c", pkg], env=env).decode().strip() == "32"
