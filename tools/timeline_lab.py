#!/usr/bin/env python
"""Wall-clock timeline of the factorisation kernels of concurrent elimination chains (tools/bin/libpgs_tl.so =
libpgs.so built with -DSKY_TIMELINE): for panels [d0, d0 + nd) of every chain, when the first CTA of diag / trsm /
next / rest started and when the last one ended (%globaltimer), one LM iteration of a BASELINE config.
  python tools/timeline_lab.py --config 3 [--chains 2]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--chains", type=int, default=0)
    args = ap.parse_args()
    import solve_keyframe_pose_graph_b200 as pgs
    from solve_keyframe_pose_graph_b200 import capi, problems
    capi.library_path = lambda: os.path.join(ROOT, "tools", "bin", "libpgs_tl.so")
    p = problems.build_problem(args.config)
    S = problems.load_into_solver(p, linear_solver=capi.SKYLINE_CHOLESKY, max_num_iterations=1, chains=args.chains)
    s = S.solve()
    L = capi.lib()
    buf = np.zeros(4 * 5 * 64 * 2, np.uint64)
    d0, nd, kinds = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.pgs_debug_timeline(buf.ctypes.data_as(C.c_void_p), C.byref(d0), C.byref(nd), C.byref(kinds))
    assert rc == 0
    d0, nd, kinds = d0.value, nd.value, kinds.value
    tl = buf[:4 * kinds * nd * 2].reshape(4, kinds, nd, 2).astype(np.int64)
    names = ["diag", "trsm", "next", "rest"]
    live = [c for c in range(4) if tl[c, 0, :, 1].max() > 0]
    t0 = min(int(tl[c, k, :, 0][tl[c, k, :, 1] > 0].min()) for c in live for k in range(4) if (tl[c, k, :, 1] > 0).any())
    print(f"config {args.config}, {s['n_chains']} chains, ms_linear_solve {s['ms_linear_solve']:.1f}; times in us from the first stamp; panels {d0}..{d0 + nd - 1}")
    for c in live:
        print(f"--- chain {c}")
        print("panel  " + "  ".join(f"{n:>6}_in {n:>6}_out" for n in names) + "   period(rest_in)")
        prev = None
        for i in range(nd):
            row = []
            for k in range(4):
                a, b = tl[c, k, i]
                row.append(f"{(a - t0) / 1e3:9.1f} {(b - t0) / 1e3:10.1f}" if b > 0 else " " * 20)
            r_in = tl[c, 3, i, 0]
            per = f"{(r_in - prev) / 1e3:8.1f}" if prev is not None and tl[c, 3, i, 1] > 0 else ""
            prev = r_in
            print(f"{d0 + i:5d}  " + "  ".join(row) + "   " + per)
        dur = {n: np.mean([(tl[c, k, i, 1] - tl[c, k, i, 0]) / 1e3 for i in range(nd) if tl[c, k, i, 1] > 0]) for k, n in enumerate(names)}
        print("mean durations (us): " + ", ".join(f"{n} {v:.1f}" for n, v in dur.items()))
    S.close()


if __name__ == "__main__":
    main()
