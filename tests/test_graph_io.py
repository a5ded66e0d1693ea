"""On-disk formats of the reference (csrc/host/GraphIO.{h,cpp}).  CPU only (dry-run facade).
Golden vector: the sample of solved_posegraph.json that the reference quotes in its own source
(src/NodeDataManager.cpp:892-908, 955-995) — real output of the reference — pins the matrix string format, the parser,
prettyprintMatrix4d / R2ypr and the layout.  Round trips pin the writers against the readers."""
import json
import os

import numpy as np
import pytest

from oracle import frontend, pgo
import solve_keyframe_pose_graph_b200 as pgs
from solve_keyframe_pose_graph_b200 import facade, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_solved_posegraph_sample.json")


def test_reference_sample_parses_prints_and_reprints_identically():
    ref = json.load(open(GOLD))
    T, st, w, sid = facade.io_load_solved_posegraph(GOLD)
    assert len(T) == 3 and list(st) == [n["stampNSec"] for n in ref["SolvedPoseGraph"]] and list(w) == [0, 0, 0] and list(sid) == [0, 0, 0]
    for i, node in enumerate(ref["SolvedPoseGraph"]):
        M = np.array([[float(x) for x in row.split(",")] for row in node["w_T_c"]["data"].split("\n")])
        assert np.array_equal(T[i], M)                                                    # parser (RawFileIO.cpp:372-409)
        assert facade.io_prettyprint(T[i]) == node["w_T_c"]["data_pretty"]               # R2ypr + %4.3f (PoseManipUtils.cpp:143-158,206-215)
        assert facade.io_mat_to_string(T[i], solved_layout=True) == node["w_T_c"]["data"]   # Eigen FullPrecision = 16 significant digits
        assert np.array_equal(facade.io_string_to_mat(facade.io_mat_to_string(T[i])), M)    # the ',' / ';' layout of the log files
        # the oracle's R2ypr agrees with the reference's printed angles
        ypr = pgo.r2ypr_deg(M)
        assert node["w_T_c"]["data_pretty"].startswith(":YPR(deg)=(%4.3f,%4.3f,%4.3f)" % tuple(ypr))
    assert facade.io_string_to_mat("1,2,3;4,5,6") is None


def test_log_posegraph_round_trip_reproduces_the_session(tmp_path):
    g = synth.generate_config(4, n_nodes=60, n_interworld=12)
    F = facade.Facade(odom_fanout=3, dry_run=True); F.ingest(g)
    assert F.solve_once()
    assert F.save_json(tmp_path) == 3                                                   # no composer pass in a dry run
    J = json.load(open(tmp_path / "log_posegraph.json"))
    assert J["meta_data"]["getNodeLen"] == g["N"] == len(J["nodes"]) and J["meta_data"]["getEdgeLen"] == len(g["la"]) == len(J["loopedges"])
    assert [n["world_id"] for n in J["nodes"]] == [F.which_world(s) for s in g["stamps"]]
    assert len(J["kidnap_info"]) == 3 and [w["nodeidx_of_world_i_ended"] for w in J["world_info"]] == [F.world_end(w) for w in range(4)]
    e0 = J["loopedges"][0]
    assert (e0["idx0"], e0["idx1"]) == (int(g["la"][0]), int(g["lb"][0])) and e0["code"] in (1, 2) and e0["weight"] == g["lw"][0]
    assert sorted(J["nodes"][0].keys()) == ["cov", "idx", "timestamp", "wTc", "wTc_pretty", "world_id"]        # the reference's keys (:517-536)
    O = json.load(open(tmp_path / "log_optimized_poses.json"))
    assert O["meta_data"]["nNodes"] == g["N"] and len(O["PoseGraphSLAM_loopedgeinfo"]) == len(g["la"])
    assert "switching_var_after_opt" in O["PoseGraphSLAM_loopedgeinfo"][0]
    # a fresh facade loaded from the file behaves like the original one
    G = facade.Facade(odom_fanout=3, dry_run=True); G.load_posegraph_json(tmp_path)
    assert G.n_keyframes() == g["N"] and G.n_worlds() == 4
    assert [G.which_world(s) for s in g["stamps"]] == [F.which_world(s) for s in g["stamps"]]
    G.n_loop = len(g["la"])
    assert G.solve_once()
    assert [G.world_setid(w) for w in range(4)] == [F.world_setid(w) for w in range(4)]
    a, b = F.odom_terms(), G.odom_terms()
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["umf"], b["umf"]) and np.allclose(a["w"], b["w"], rtol=1e-12)
    assert np.allclose(a["t"], b["t"], atol=1e-12)                                      # 16 significant digits survive the text round trip
    qa, ta = F.poses(); qb, tb = G.poses()
    assert np.allclose(ta, tb, atol=1e-9)
    # WorldsData of solved_posegraph.json (Worlds.cpp:442-497): relative poses, world stamps and the union-find op-log,
    # replayed by an empty facade (Worlds::loadStateFromDisk, :499-640)
    W = json.load(open(tmp_path / "solved_posegraph.json"))["WorldsData"]
    assert W["disjoint_set"]["log_string"].startswith("add_element:0;add_element:1;add_element:2;add_element:3;") and "union_sets:" in W["disjoint_set"]["log_string"]
    assert len(W["vec_world_starts"]) == 4 and len(W["vec_world_ends"]) == 3 and len(W["rel_pose_between_worlds__wb_T_wa"]) >= 3
    H = facade.Facade(dry_run=True); H.load_worlds_state(tmp_path / "solved_posegraph.json")
    assert [H.world_setid(w) for w in range(5)] == [F.world_setid(w) for w in range(4)] + [-1]   # four worlds known to the union-find, same roots
    for m_ in range(4):
        for n_ in range(4):
            assert np.allclose(H.pose_between_worlds(m_, n_), F.pose_between_worlds(m_, n_), atol=1e-9)
    F.close(); G.close(); H.close()


@pytest.mark.gpu
def test_solved_posegraph_written_after_a_device_compose(tmp_path):
    g = synth.generate_config(4, n_nodes=80, n_worlds=3, n_interworld=10)
    F = facade.Facade(odom_fanout=3); F.ingest(g)
    assert F.solve_once()
    T, wid = F.compose()
    assert F.save_json(tmp_path) == 7
    T2, st, w2, sid = facade.io_load_solved_posegraph(tmp_path / "solved_posegraph.json")
    assert np.allclose(T2, T, rtol=0, atol=1e-12 * max(1.0, np.abs(T).max())) and np.array_equal(w2, wid) and np.array_equal(st, g["stamps"])
    J = json.load(open(tmp_path / "solved_posegraph.json"))
    assert sorted(J["SolvedPoseGraph"][0].keys()) == ["seq", "setID_of_worldID", "stampNSec", "w_T_c", "worldID"]
    assert len(J["KidnapTimestamps"]["kidnap_starts"]) == 2 == len(J["KidnapTimestamps"]["kidnap_ends"])
    # WorldsData (Worlds.cpp:442-497): relative poses, world stamps and the union-find op-log, replayed by a fresh facade
    W = J["WorldsData"]
    assert W["disjoint_set"]["log_string"].startswith("add_element:0;add_element:1;add_element:2;") and "union_sets:" in W["disjoint_set"]["log_string"]
    assert len(W["vec_world_starts"]) == 3 and len(W["vec_world_ends"]) == 2 and len(W["rel_pose_between_worlds__wb_T_wa"]) >= 2
    G = facade.Facade(dry_run=True); G.load_worlds_state(tmp_path / "solved_posegraph.json")
    assert [G.world_setid(w) for w in range(4)] == [F.world_setid(w) for w in range(3)] + [-1]
    for m_ in range(3):
        for n_ in range(3):
            assert np.allclose(G.pose_between_worlds(m_, n_), F.pose_between_worlds(m_, n_), atol=1e-9)
    F.close(); G.close()


def test_worlds_op_log_example_from_the_reference_replays(tmp_path):
    # the example the reference gives for the union-find op-log (src/Worlds.cpp:549-554)
    state = {"WorldsData": {"rel_pose_between_worlds__wb_T_wa": [], "vec_world_starts": [{"stampNSec": 10}, {"stampNSec": 20}, {"stampNSec": 30}],
                            "vec_world_ends": [{"stampNSec": 15}, {"stampNSec": 25}],
                            "disjoint_set": {"debug_string": "\\t\\t\\tadd_element( 0)\\n\\t\\t\\tadd_element( 1)\\n\\t\\t\\tadd_element( 2)\\n\\t\\t\\tunion_sets( 0,2)\\n",
                                             "log_string": "add_element:0;add_element:1;add_element:2;union_sets:0,2;"}}}
    f = tmp_path / "solved_posegraph.json"; f.write_text(json.dumps(state))
    F = facade.Facade(dry_run=True); F.load_worlds_state(f)
    # union_sets(max, min) with link-by-rank: on the tie world 0 stays the root (src/Worlds.cpp:168, DisjointSet.h:241-257)
    assert [F.world_setid(w) for w in range(4)] == [0, 1, 0, -1]
    bad = dict(state); bad["WorldsData"] = dict(state["WorldsData"], disjoint_set={"debug_string": "", "log_string": "add_element:0;merge:0,1;"})
    g = tmp_path / "bad.json"; g.write_text(json.dumps(bad))
    G = facade.Facade(dry_run=True)
    with pytest.raises(pgs.PgsError, match="unknown op"):
        G.load_worlds_state(g)
    F.close(); G.close()


def test_pose_covariance_round_trips_through_log_posegraph(tmp_path):
    # nav_msgs/Odometry pose.covariance is kept with the keyframe (NodeDataManager.cpp:55-63) and written as a 6x6 string
    rng = np.random.default_rng(2)
    F = facade.Facade(dry_run=True)
    covs = []
    for i in range(5):
        A = rng.normal(size=(6, 6)); cov = A @ A.T; covs.append(cov)
        F.camera_pose_callback(10**9 + i * 10**8, [float(i), 0.0, 0.0], [0, 0, 0, 1.0], cov)
    F.save_json(tmp_path)
    J = json.load(open(tmp_path / "log_posegraph.json"))
    for i, n in enumerate(J["nodes"]):
        M = np.array([[float(x) for x in row.split(",")] for row in n["cov"].split(";")])
        assert M.shape == (6, 6) and np.allclose(M, covs[i], rtol=1e-15, atol=0)
    G = facade.Facade(dry_run=True); G.load_posegraph_json(tmp_path); G.save_json(tmp_path / "..")
    J2 = json.load(open(tmp_path / ".." / "log_posegraph.json"))
    assert [n["cov"] for n in J2["nodes"]] == [n["cov"] for n in J["nodes"]]
    F.close(); G.close()


def test_mutated_state_files_are_rejected_or_loaded_but_never_crash(tmp_path):
    """The three JSON files come from disk: byte-level damage (flips, cuts, duplicated and inserted fragments) must end in
    a PgsError or in a session that still saves, never in a crash.  (The same loop ran 750 file sets under ASan/UBSan.)"""
    import random
    import shutil
    rnd = random.Random(7)
    g = synth.generate_config(4, n_nodes=20, n_interworld=6)
    F = facade.Facade(odom_fanout=2, dry_run=True); F.ingest(g); assert F.solve_once(); F.save_json(tmp_path); F.close()
    files = {f: open(tmp_path / f, "rb").read() for f in ("log_posegraph.json", "log_optimized_poses.json", "solved_posegraph.json")}

    def mutate(b):
        b = bytearray(b)
        for _ in range(rnd.randint(1, 6)):
            op, i = rnd.randint(0, 4), rnd.randrange(len(b))
            if op == 0: b[i] = rnd.randrange(256)
            elif op == 1: del b[i:i + rnd.randint(1, 40)]
            elif op == 2: b[i:i] = bytes(rnd.choice(b'{}[]",:0123456789.-eE;\\ntrufalsn') for _ in range(rnd.randint(1, 12)))
            elif op == 3: b = b[:i]
            else:
                j = rnd.randrange(len(b)); b[i:i] = b[j:j + rnd.randint(1, 60)]
            if not b: b = bytearray(b"{")
        return bytes(b)

    accepted = rejected = 0
    for it in range(40):
        d = tmp_path / f"m{it}"; d.mkdir()
        for f, v in files.items():
            open(d / f, "wb").write(mutate(v) if rnd.random() < 0.7 else v)
        G = facade.Facade(odom_fanout=2, dry_run=True)
        for call in (lambda: G.load_posegraph_json(d), lambda: G.load_worlds_state(d / "solved_posegraph.json"), lambda: G.load_state(),
                     lambda: G.solve_once(True), lambda: G.save_json(d), lambda: facade.io_load_solved_posegraph(d / "solved_posegraph.json")):
            try:
                call(); accepted += 1
            except pgs.PgsError:
                rejected += 1
        G.close(); shutil.rmtree(d)
    assert accepted > 0 and rejected > 0


def test_replay_tool_reads_a_recorded_run_and_compares_with_its_recorded_solution(tmp_path):
    """tests/replay_reference_run.py on a directory laid out as the reference leaves it (log_posegraph.json +
    log_optimized_poses.json).  The recorded solution here is the oracle's own, written in the reference's format, so the
    CPU replay must land on it: this checks the readers, the block construction from a recorded graph and the comparison,
    which is what a real recorded run would go through."""
    import replay_reference_run as rr
    from oracle import frontend, pgo
    g = synth.generate_config(2, n_nodes=150, n_loop=20)
    F = facade.Facade(dry_run=True); F.ingest(g); assert F.solve_once(); F.save_json(tmp_path); F.close()
    M = frontend.Manager(); M.ingest(g)
    R = frontend.ReferenceFrontEnd(M, odom_fanout=5); s = R.trigger(solve=True)
    J = json.load(open(tmp_path / "log_optimized_poses.json"))
    assert len(J["PoseGraphSLAM_nodes"]) == 150
    for i, node in enumerate(J["PoseGraphSLAM_nodes"]):
        node["wTc_opt"] = facade.io_mat_to_string(pgo.pose_to_mat4(R.opt_q[i], R.opt_t[i]))
    for e in J["PoseGraphSLAM_loopedgeinfo"]:
        e["switching_var_after_opt"] = R.opt_s[e["getEdge_i"]]
    json.dump(J, open(tmp_path / "log_optimized_poses.json", "w"))
    out = rr.replay(str(tmp_path), use_oracle=True, fanout=5)
    assert out["nodes"] == 150 and out["loop_edges"] == 20 and out["blocks"]["odometry"] == len(R.odom) and out["blocks"]["regularisers"] == 1
    assert out["translation_dev_m"]["max"] < 1e-9 and out["rotation_dev_rad"]["max"] < 1e-7
    assert out["switches"] == dict(compared=20, same_state=20)
    assert abs(out["cost_of_reference_solution"] - s["final_cost"]) <= 1e-9 * s["final_cost"] and out["cost_at_odometry"] > s["final_cost"]
    assert abs(out["final_cost"] - s["final_cost"]) <= 1e-12 * s["final_cost"]
