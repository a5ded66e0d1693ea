"""The drop-in demonstrated with the reference's own code (needs a GPU and oracle/_ref/libref_frontend.so): the reference's NodeDataManager / Worlds / PoseGraphSLAM sources run unmodified
(tests/test_reference_frontend.py explains how) and every ceres::Solve they issue is served by libpgs.so — the product's
raw solver C-ABI — instead of Ceres.  A second reference instance gets the same session served by the CPU oracle; after every
wake-up the two must agree within north_star's tolerances (1e-5 m, 1e-4 rad, same switch states, cost 1e-5 relative).

    python tests/reference_node_with_libpgs.py        -> one line per wake-up, exit status 1 on disagreement
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import solve_keyframe_pose_graph_b200 as pgs  # noqa: E402
from oracle import pgo  # noqa: E402
from solve_keyframe_pose_graph_b200 import synth  # noqa: E402
from test_reference_frontend import ReferenceNode, dp  # noqa: E402


def poses_of(Ms):
    qt = [pgo.mat4_to_pose(X) for X in Ms]
    return np.array([x[0] for x in qt]).reshape(-1, 4), np.array([x[1] for x in qt]).reshape(-1, 3)


def server(R, backend, log):
    def serve():
        B = R.blocks(); q, t, s, _ = R.variables()
        od, lo, rg = B["type"] == 0, B["type"] == 1, B["type"] == 2
        oq, ot = poses_of(B["obs"][od]); lq, lt = poses_of(B["obs"][lo]); rq, rt = poses_of(B["obs"][rg])
        if backend == "libpgs":
            S = pgs.PoseGraphSolver()
            S.set_nodes(q, t); S.add_odom_edges(B["c1"][od], B["c2"][od], oq, ot, B["w"][od])
            S.add_loop_edges(B["c2"][lo], B["c1"][lo], lq, lt, B["w"][lo])          # (a, b) with b_T_a: bound as (b, a, switch), as the reference binds it
            S.set_switches(s[B["sw"][lo]]); S.set_regularizers(B["c1"][rg], rq, rt, B["w"][rg])
            out = S.solve(); q2, t2 = S.poses(); sw = S.switches(); S.close()
        else:
            S = pgo.Problem(); S.set_nodes(q, t); S.add_odom_edges(B["c1"][od], B["c2"][od], oq, ot, B["w"][od])
            S.add_loop_edges(B["c1"][lo], B["c2"][lo], lq, lt, B["w"][lo], s_init=s[B["sw"][lo]]); S.set_regularizers(B["c1"][rg], rq, rt, B["w"][rg])
            out = S.solve(); q2, t2 = S.poses(); sw = S.switches()
        s2 = s.copy(); s2[B["sw"][lo]] = sw
        q2, t2, s2 = (np.ascontiguousarray(x, dtype=np.float64) for x in (q2, t2, s2))
        R.L.refslam_write_vars(R.h, q2.ctypes.data_as(dp), t2.ctypes.data_as(dp), s2.ctypes.data_as(dp))
        log.append(out)
    return serve


def run_backend(backend, out_path):
    """One reference instance per process (two in one process share the C++ runtime's unique symbols and trip over each other)."""
    g = synth.generate_config(2, n_nodes=600, n_loop=90)
    order = np.argsort(np.maximum(g["la"], g["lb"]), kind="stable")
    R = ReferenceNode(); log = []
    R.L.refslam_set_solve_callback.argtypes = [C.c_void_p, C.c_void_p]; R.L.refslam_write_vars.argtypes = [C.c_void_p, dp, dp, dp]
    cb = C.CFUNCTYPE(None)(server(R, backend, log))
    R.L.refslam_set_solve_callback(R.h, C.cast(cb, C.c_void_p))
    res = {}; epos = 0
    try:
        for k, lo in enumerate(range(0, 600, 200)):
            take = []
            while epos < len(order) and max(g["la"][order[epos]], g["lb"][order[epos]]) < lo + 200:
                take.append(order[epos]); epos += 1
            take = np.array(take, dtype=int)
            R.add_nodes(g["stamps"][lo:lo + 200], g["q"][lo:lo + 200], g["t"][lo:lo + 200])
            R.add_loop_edges(g["stamps"], g["la"][take], g["lb"][take], g["lq"][take], g["lt"][take], g["lw"][take])
            assert R.wakeup()
            q, t, s, _ = R.variables()
            res.update({f"q{k}": q, f"t{k}": t, f"s{k}": s, f"cost{k}": log[-1]["final_cost"], f"iters{k}": len(log[-1]["iterations"]), f"edges{k}": epos})
    finally:
        R.close()
    np.savez(out_path, **res)


def main():
    import subprocess
    import tempfile
    if len(sys.argv) == 3:
        return run_backend(sys.argv[1], sys.argv[2])
    d = tempfile.mkdtemp(prefix="refnode_")
    out = {}
    for backend in ("libpgs", "oracle"):
        path = os.path.join(d, backend + ".npz")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), backend, path], capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print(r.stdout[-1500:] + r.stderr[-3000:]); sys.exit(2)
        out[backend] = np.load(path)
    a, b = out["libpgs"], out["oracle"]
    bad = 0
    for k in range(3):
        dt = np.abs(a[f"t{k}"] - b[f"t{k}"]).max(); dr = (2 * np.arccos(np.abs(np.sum(a[f"q{k}"] * b[f"q{k}"], axis=1)).clip(0, 1))).max()
        ca, cb_ = float(a[f"cost{k}"]), float(b[f"cost{k}"])
        dc = abs(ca - cb_) / cb_
        ok = dt < 1e-5 and dr < 1e-4 and np.array_equal(a[f"s{k}"] > 0.5, b[f"s{k}"] > 0.5) and dc < 1e-5 and int(a[f"iters{k}"]) == int(b[f"iters{k}"])
        bad += not ok
        print(f"wake-up at {200 * (k + 1)} keyframes / {int(a[f'edges{k}'])} loop edges: dt {dt:.2e} m  drot {dr:.2e} rad  cost {ca:.6g} vs {cb_:.6g} ({dc:.1e})  "
              f"iterations {int(a[f'iters{k}']) - 1}/{int(b[f'iters{k}']) - 1}  {'ok' if ok else 'MISMATCH'}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
