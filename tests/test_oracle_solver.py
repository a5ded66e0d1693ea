"""Pins for the oracle's restatement of ceres::Solve: the sparse linear
step against a dense full-system solve in numpy, trust-region behaviour, switch behaviour, and the
committed golden trajectories."""
import os

import numpy as np
import pytest

from make_golden import CASES, run_case
from oracle import pgo
from util_graphs import load_oracle, random_graph

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dense_lm_step(g, ev, radius):
    """(J~^T J~ + D^2) y = J~^T r over ALL unknowns (6N poses + switches), Jacobi-scaled, dense numpy."""
    N, El = g["N"], len(g["la"])
    n = 6 * N + El
    rows = []
    def block(J, r, cols):
        for i in range(J.shape[0]):
            row = np.zeros(n + 1); row[cols] = J[i]; row[n] = r[i]; rows.append(row)
    for e in range(len(g["oc1"])):
        c1, c2 = g["oc1"][e], g["oc2"][e]
        block(ev["J_o"][e], ev["r_o"][e], np.r_[6 * c1:6 * c1 + 6, 6 * c2:6 * c2 + 6])
    for e in range(El):
        c1, c2 = g["lb"][e], g["la"][e]
        block(ev["J_l"][e], ev["r_l"][e], np.r_[6 * c1:6 * c1 + 6, 6 * c2:6 * c2 + 6, 6 * N + e])
    for k in range(len(g["rn"])):
        i = g["rn"][k]; block(ev["J_r"][k], ev["r_r"][k], np.r_[6 * i:6 * i + 6])
    A = np.array(rows); J, r = A[:, :n], A[:, n]
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    Js = J * scale
    diag = np.clip((Js * Js).sum(axis=0), 1e-6, 1e32)
    H = Js.T @ Js + np.diag(diag / radius)
    y = np.linalg.solve(H, Js.T @ r)
    step = -y
    m = Js @ step
    return (step * scale), -m @ (r + m / 2)


@pytest.mark.parametrize("radius", [1e4, 2.5e7, 3.0])
def test_linear_step_equals_dense_full_system_solve(radius):
    g = random_graph(40, 3, 10, outlier_frac=0.2, seed=21)
    P = load_oracle(g)
    ev = P.evaluate(autodiff=True)
    delta, mcc = dense_lm_step(g, ev, radius)
    dp, ds, m = P.linear_step(radius)
    assert np.abs(dp.ravel() - delta[: 6 * g["N"]]).max() < 1e-8 * max(1, np.abs(delta).max())
    assert np.abs(ds - delta[6 * g["N"]:]).max() < 1e-9
    assert abs(m - mcc) < 1e-8 * max(1, abs(mcc))


def test_autodiff_and_closed_form_lm_agree():
    g = random_graph(120, 3, 30, outlier_frac=0.1, seed=22)
    A = load_oracle(g); B = load_oracle(g)
    sa = A.solve(pgo.default_options(use_autodiff=1)); sb = B.solve(pgo.default_options(use_autodiff=0, num_threads=4))
    assert sa["termination"] == sb["termination"] and len(sa["iterations"]) == len(sb["iterations"])
    assert abs(sa["final_cost"] - sb["final_cost"]) < 1e-9 * sa["final_cost"]
    assert np.abs(A.poses()[1] - B.poses()[1]).max() < 1e-7


def test_trust_region_bookkeeping():
    g = random_graph(200, 3, 50, outlier_frac=0.1, seed=23)
    P = load_oracle(g); s = P.solve()
    it = s["iterations"]
    assert it[0]["iteration"] == 0 and it[0]["trust_region_radius"] == 1e4
    assert len(it) <= 11 and s["termination"] in ("CONVERGENCE", "NO_CONVERGENCE")
    cost = it[0]["cost"]; radius = 1e4
    for r in it[1:]:
        if r["step_is_successful"]:
            assert r["cost"] < cost and r["relative_decrease"] > 1e-3
            radius = min(1e16, radius / max(1 / 3, 1 - (2 * r["relative_decrease"] - 1) ** 3)); cost = r["cost"]
            assert np.isclose(r["trust_region_radius"], radius, rtol=1e-12)
    assert np.isclose(s["final_cost"], cost)


def test_switch_cliff_and_outlier_rejection():
    # gross outliers (||e||^2 >> 1/8) slide to s = 0, tight inliers stay near 1
    g = random_graph(300, 3, 60, outlier_frac=0.15, seed=24)
    assert g["lout"].sum() >= 3
    P = load_oracle(g)
    s = P.solve(pgo.default_options(max_num_iterations=50, function_tolerance=1e-12))
    sw = P.switches()
    assert np.all(np.abs(sw[g["lout"]]) < 1e-3) and np.all(sw[~g["lout"]] > 0.95)
    ev = P.evaluate()
    assert ev["cost"] < 0.5 * s["initial_cost"]


def test_tight_convergence_reaches_a_stationary_point():
    g = random_graph(80, 3, 20, seed=25)
    P = load_oracle(g)
    s = P.solve(pgo.default_options(max_num_iterations=200, function_tolerance=1e-16, parameter_tolerance=1e-14, gradient_tolerance=1e-12))
    ev = P.evaluate()
    assert np.abs(ev["g_p"]).max() < 1e-6 and np.abs(ev["g_s"]).max() < 1e-6, s["termination"]


def test_consistent_graph_terminates_by_gradient_tolerance_at_iteration_zero():
    g = random_graph(30, 2, 0, seed=26, noise=0.0, reg=True)
    # noise-free odometry: odometry edges are exactly satisfied by the initial guess, regulariser at anchor
    P = load_oracle(g); s = P.solve()
    assert s["initial_cost"] < 1e-20 and s["termination"] == "CONVERGENCE" and len(s["iterations"]) == 1


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_committed_golden(name):
    G = np.load(os.path.join(GOLD, name + ".npz"))
    d = run_case(CASES[name])
    assert d["termination"] == str(G["termination"])
    assert np.allclose(d["iter_cost"], G["iter_cost"], rtol=1e-9) and np.array_equal(d["iter_success"], G["iter_success"])
    assert np.allclose(d["iter_radius"], G["iter_radius"], rtol=1e-9)
    assert np.abs(d["t"] - G["t"]).max() < 1e-7 and np.abs(d["switches"] - G["switches"]).max() < 1e-7
    assert np.allclose(d["J_o_head"], G["J_o_head"], atol=1e-12) and np.allclose(d["J_l_head"], G["J_l_head"], atol=1e-12)


def _scipy_residual_vector(g):
    """The whole residual vector of the reference's problem (CeresResidues.h:47-66,105-127,186-198) written with scipy
    rotations and 4x4 matrices only: nothing of the oracle's code is involved.  x = [q (4N), t (3N), s (El)]; the
    quaternions are normalised inside, so the ambient problem has the manifold problem's stationary points."""
    from scipy.spatial.transform import Rotation as Rot
    N, El = g["N"], len(g["la"])

    def edge(q1, t1, q2, t2, oq, ot):
        R1, R2, Ro = Rot.from_quat(q1), Rot.from_quat(q2), Rot.from_quat(oq)
        R12 = R1.inv() * R2
        dq = (R12.inv() * Ro).as_quat()
        dq = dq if dq[3] >= 0 else -dq                           # the sign Hamilton products give near the solution
        return np.r_[R12.inv().apply(ot - R1.inv().apply(t2 - t1)), 2 * dq[:3]]

    def T44(q, t):
        T = np.eye(4); T[:3, :3] = Rot.from_quat(q).as_matrix(); T[:3, 3] = t; return T

    def fun(x):
        q = x[:4 * N].reshape(N, 4); q = q / np.linalg.norm(q, axis=1, keepdims=True)
        t = x[4 * N:7 * N].reshape(N, 3); s = x[7 * N:]
        out = []
        for e in range(len(g["oc1"])):
            a, b = g["oc1"][e], g["oc2"][e]
            out.append(g["ow"][e] * edge(q[a], t[a], q[b], t[b], g["oq"][e], g["ot"][e]))
        for e in range(El):
            c1, c2 = g["lb"][e], g["la"][e]                        # loop edge (a,b) is bound as (b,a), PoseGraphSLAM.cpp:1553-1554
            out.append(s[e] * np.r_[edge(q[c1], t[c1], q[c2], t[c2], g["lq"][e], g["lt"][e]), 1.0 - s[e]])
        for k in range(len(g["rn"])):
            D = np.linalg.inv(T44(g["rq"][k], g["rt"][k])) @ T44(q[g["rn"][k]], t[g["rn"][k]])
            dq = Rot.from_matrix(D[:3, :3]).as_quat(); dq = dq if dq[3] >= 0 else -dq
            out.append(g["rw"][k] * np.r_[D[:3, 3], 2 * dq[:3]])
        return np.concatenate(out)
    return fun


def test_minimiser_agrees_with_an_independent_scipy_least_squares_solve():
    """Parity pin that involves neither the oracle's functors nor its LM: MINPACK's Levenberg-Marquardt on a scipy
    restatement of the residuals, started at the same point, must end in the minimiser the oracle finds under tight
    options: same cost to 1e-9 relative, poses within north_star's 1e-5 m / 1e-4 rad (MINPACK differentiates by finite
    differences, so it stalls a little short along the weak whole-trajectory mode), switches to 1e-5."""
    from scipy.optimize import least_squares
    from util_graphs import rot_angle_between
    g = random_graph(14, 2, 3, seed=31)
    N, El = g["N"], len(g["la"])
    fun = _scipy_residual_vector(g)
    x0 = np.r_[g["q"].ravel(), g["t"].ravel(), np.full(El, 0.99)]
    P = load_oracle(g)
    assert abs(0.5 * np.sum(fun(x0) ** 2) - P.evaluate(jac=False)["cost"]) <= 1e-12 * max(1.0, P.evaluate(jac=False)["cost"])
    sol = least_squares(fun, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=20000)
    s = P.solve(pgo.default_options(max_num_iterations=300, function_tolerance=1e-16, parameter_tolerance=1e-15, gradient_tolerance=1e-13))
    assert abs(sol.cost - s["final_cost"]) <= 1e-9 * s["final_cost"], (sol.cost, s["final_cost"], s["termination"])
    q = sol.x[:4 * N].reshape(N, 4); q = q / np.linalg.norm(q, axis=1, keepdims=True)
    t = sol.x[4 * N:7 * N].reshape(N, 3)
    qo, to = P.poses()
    assert np.abs(t - to).max() < 1e-5
    assert rot_angle_between(q, qo).max() < 1e-4
    assert np.abs(sol.x[7 * N:] - P.switches()).max() < 1e-5


def test_rejected_steps_leave_the_residuals_of_the_current_point_in_place():
    """Ceres evaluates the candidate of an LM step with residuals == NULL, so after a rejected step the next step is
    again built from J(x)^T r(x).  (Until round 2 the oracle's cost-only evaluation overwrote r with the candidate's
    residuals; after a rejection the following steps then blew up — candidate costs 1e8, 1e22, 1e50, ... in
    config3_small — which the CUDA solve, comparing against the golden, brought to light.)  With the radius halved,
    quartered, ... after every rejection, the rejected candidates have to come back towards the current cost."""
    G = np.load(os.path.join(GOLD, "config3_small.npz"))
    ok = G["iter_success"].astype(bool); cost = G["iter_cost"]
    assert (~ok).sum() >= 5 and ok[-1]                                  # five rejections in a row, then progress again
    x_cost = cost[2]
    rejected = cost[3:8]
    assert np.all(rejected < 10 * x_cost) and rejected[-1] < rejected[0]
    assert cost[-1] < 0.1 * x_cost


def test_fixed_cost_follows_ceres_reduced_program():
    """Residual blocks whose parameter blocks are all constant leave the reduced program: Summary::fixed_cost holds their
    cost, the iteration table is without it, initial/final cost include it (ceres/solver.h, Program::RemoveFixedBlocks)."""
    g = random_graph(80, 2, 10, seed=33)
    P = load_oracle(g); P.set_constant_nodes(0, 30)
    e = P.evaluate(autodiff=True)
    both = (g["oc1"] < 30) & (g["oc2"] < 30)
    fc = 0.5 * (np.sum(e["r_o"][both] ** 2) + np.sum(e["r_r"] ** 2))          # the regulariser sits on keyframe 0
    s = P.solve()
    assert fc > 0 and abs(s["fixed_cost"] - fc) <= 1e-13 * fc
    assert abs(s["initial_cost"] - e["cost"]) <= 1e-13 * e["cost"] and abs(s["iterations"][0]["cost"] - (e["cost"] - fc)) <= 1e-12 * e["cost"]
    assert abs(s["final_cost"] - (s["iterations"][-1]["cost"] + fc)) <= 1e-12 * s["final_cost"] or not s["iterations"][-1]["step_is_successful"]
    Q = load_oracle(g); assert Q.solve()["fixed_cost"] == 0.0
