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
