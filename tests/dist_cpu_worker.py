"""World-size-2 (gloo, CPU) check of the multi-GPU scheme's host logic and algebra.

Each rank takes the partition from libpgs' host-only pgs_partition, builds its partial normal equations from
the ORACLE's per-block residuals/Jacobians of the blocks it owns, eliminates its interior (poses + the switches
of its loop edges) on the locally scaled system, all-reduces [border Schur | border rhs | border diag(J^T J)]
over gloo, applies the border scaling/damping to the sum, solves the border, back-substitutes, and compares the
assembled step with the oracle's full-system LM step.  This is DESIGN.md §4 in numpy; the CUDA path is checked
against the same oracle on real GPUs by tests/test_dist_gpu.py.  Launched by tests/test_dist_cpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from util_graphs import load_oracle, random_graph  # noqa: E402
import solve_keyframe_pose_graph_b200 as pgs  # noqa: E402

LO, HI = 1e-6, 1e32


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for seed, n, nl, radius in [(3, 180, 40, 1e4), (4, 240, 70, 37.0)]:
        g = random_graph(n, 3, nl, outlier_frac=0.15, seed=seed)
        N, El = g["N"], len(g["la"])
        P = pgs.partition(N, world, g["oc1"], g["oc2"], g["la"], g["lb"], g["rn"])
        O = load_oracle(g)
        ev = O.evaluate(autodiff=True)
        nu = 6 * N + El
        H = np.zeros((nu, nu)); gr = np.zeros(nu)

        def add(J, r, idx):
            H[np.ix_(idx, idx)] += J.T @ J; gr[idx] += J.T @ r
        for e in np.nonzero(P["odom_owner"] == rank)[0]:
            c1, c2 = g["oc1"][e], g["oc2"][e]
            add(ev["J_o"][e], ev["r_o"][e], np.r_[6 * c1:6 * c1 + 6, 6 * c2:6 * c2 + 6])
        for e in np.nonzero(P["loop_owner"] == rank)[0]:
            c1, c2 = g["lb"][e], g["la"][e]     # bound as (b, a, s)
            add(ev["J_l"][e], ev["r_l"][e], np.r_[6 * c1:6 * c1 + 6, 6 * c2:6 * c2 + 6, 6 * N + e])
        for k in np.nonzero(P["reg_owner"] == rank)[0]:
            i = g["rn"][k]
            add(ev["J_r"][k], ev["r_r"][k], np.r_[6 * i:6 * i + 6])
        border_nodes = np.nonzero(P["node_owner"] < 0)[0]
        mine = np.nonzero(P["node_owner"] == rank)[0]
        b_idx = (6 * border_nodes[:, None] + np.arange(6)).ravel()
        i_idx = np.r_[(6 * mine[:, None] + np.arange(6)).ravel(), 6 * N + np.nonzero(P["loop_owner"] == rank)[0]].astype(int)
        # every unknown this rank touches is interior to it or border
        touched = np.nonzero(np.abs(H).sum(axis=1) > 0)[0]
        assert set(touched) <= set(i_idx) | set(b_idx), "an owned block reaches another rank's interior"
        d = np.diag(H)
        s_i = 1.0 / (1.0 + np.sqrt(d[i_idx]))
        D_i = np.clip(d[i_idx] * s_i ** 2, LO, HI) / radius
        Aii = s_i[:, None] * H[np.ix_(i_idx, i_idx)] * s_i[None, :] + np.diag(D_i)
        unused = d[i_idx] == 0
        Aii[unused, unused] = 1.0
        Abi = H[np.ix_(b_idx, i_idx)] * s_i[None, :]
        bi, bb = s_i * gr[i_idx], gr[b_idx]
        X = np.linalg.solve(Aii, np.c_[Abi.T, bi])
        S = H[np.ix_(b_idx, b_idx)] - Abi @ X[:, :-1]
        rhs = bb - Abi @ X[:, -1]
        buf = torch.from_numpy(np.r_[S.ravel(), rhs, d[b_idx]])
        dist.all_reduce(buf)                                   # the one border all-reduce per linear solve
        nb = len(b_idx)
        S = buf[:nb * nb].numpy().reshape(nb, nb); rhs = buf[nb * nb:nb * nb + nb].numpy(); dH = buf[nb * nb + nb:].numpy()
        s_b = 1.0 / (1.0 + np.sqrt(dH))
        S = S + np.diag(np.clip(dH * s_b ** 2, LO, HI) / (radius * s_b ** 2))
        z_b = np.linalg.solve(S, rhs)
        z_i = np.linalg.solve(Aii, bi - Abi.T @ z_b)
        step = np.zeros(nu)
        step[i_idx] = -s_i * z_i
        if rank == 0:
            step[b_idx] = -z_b
        t = torch.from_numpy(step); dist.all_reduce(t)         # gather
        dpo, dso, _ = O.linear_step(radius)
        ref = np.r_[dpo.ravel(), dso]
        err = np.abs(t.numpy() - ref).max() / max(1.0, np.abs(ref).max())
        assert err < 1e-8, (seed, err)
        # ownership: every block has exactly one owner in [0, world)
        for k in ("odom_owner", "loop_owner", "reg_owner"):
            assert ((P[k] >= 0) & (P[k] < world)).all()
        if rank == 0:
            print(f"dist-cpu seed {seed}: N={N} border={len(border_nodes)} step err {err:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
