"""The reference's alternative edge functors (src/CeresResidues.h:226-546, switched off in its build):
known-answer tests and finite differences pin the oracle restatement (oracle/pgo_fourdof.hpp); the arithmetic of the
device kernel (csrc/pgs_fourdof.cuh, compiled for the host by tests/fourdof_hostcheck.cpp) and, on a GPU, the kernel itself
through the C-ABI (include/pgs_fourdof.h) are compared with it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from oracle import pgo

HERE = os.path.dirname(os.path.abspath(__file__))
I4 = np.array([0.0, 0.0, 0.0, 1.0])
Z3 = np.zeros(3)


def one(kind, rot1, t1, rot2, t2, obs_rot, obs_t, w=1.0, s=None):
    out = pgo.fourdof_eval(kind, np.array([rot1, rot2]), np.array([t1, t2]), [0], [1], [obs_rot], [obs_t],
                           weight=[w], sw=None if s is None else [s])
    return out["r"][0], out["J"][0]


def random_edges(kind, n_nodes, n_edges, seed):
    """Random poses, random node pairs, observation = true relative pose x moderate noise (|angle| <~ 0.6 rad, so the
    yaw/pitch/roll extraction stays clear of its pitch = +-90 deg singularity)."""
    rng = np.random.default_rng(seed)
    t = rng.normal(size=(n_nodes, 3)) * 5
    c1 = rng.integers(0, n_nodes, n_edges).astype(np.int32)
    c2 = ((c1 + rng.integers(1, n_nodes, n_edges)) % n_nodes).astype(np.int32)
    if kind == 2:
        ypr = np.c_[rng.uniform(-180, 180, n_nodes), rng.uniform(-20, 20, n_nodes), rng.uniform(-20, 20, n_nodes)]
        obs_t = rng.normal(size=(n_edges, 3)) * 3
        rel = ypr[c2, 0] - ypr[c1, 0] + rng.normal(size=n_edges) * 5           # unwrapped on purpose: NormalizeAngle has work to do
        obs_rot = np.c_[rel, ypr[c1, 1], ypr[c1, 2]]
        return dict(rot=ypr, t=t, c1=c1, c2=c2, obs_rot=obs_rot, obs_t=obs_t, weight=None, sw=None)
    R = Rot.random(n_nodes, random_state=seed)
    q = R.as_quat()
    q[::3] *= -1                                                               # both hemispheres
    rel = R[c1].inv() * R[c2]
    noise = Rot.from_rotvec(rng.normal(size=(n_edges, 3)) * 0.2)
    obs_rot = (rel * noise).as_quat()
    obs_t = R[c1].inv().apply(t[c2] - t[c1]) + rng.normal(size=(n_edges, 3)) * 0.3
    weight = rng.uniform(0.2, 1.5, n_edges)
    sw = rng.uniform(-0.2, 1.2, n_edges) if kind == 1 else None
    return dict(rot=q, t=t, c1=c1, c2=c2, obs_rot=obs_rot, obs_t=obs_t, weight=weight, sw=sw)


def close(a, b, tol=1e-11):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


# ------------------------------------------------------------------------------------------------ known answers
def test_fourdof_identity_and_single_axis_known_answers():
    r, _ = one(0, I4, Z3, I4, Z3, I4, Z3, w=0.7)
    assert np.abs(r).max() < 1e-15
    # poses equal, observation = pure yaw / pitch / roll of 30 / 10 / -20 degrees: delta_q = q_obs, residual = w * (4, 10, 10) * angle
    for axis, deg, row, k in (("z", 30.0, 3, 4.0), ("y", 10.0, 4, 10.0), ("x", -20.0, 5, 10.0)):
        r, _ = one(0, I4, Z3, I4, Z3, Rot.from_euler(axis, deg, degrees=True).as_quat(), Z3, w=0.5)
        want = np.zeros(6); want[row] = 0.5 * k * deg
        assert np.abs(r - want).max() < 1e-12, (axis, r)
    # pure translation: c2 one metre ahead, observation says two -> delta_t = R12^T (t_obs - p12) = (1, 0, 0)
    r, _ = one(0, I4, Z3, I4, [1.0, 0, 0], I4, [2.0, 0, 0], w=0.5)
    assert np.allclose(r, [0.5, 0, 0, 0, 0, 0], atol=1e-15)


def test_fourdof_matches_a_scipy_restatement():
    g = random_edges(0, 40, 200, seed=3)
    out = pgo.fourdof_eval(0, g["rot"], g["t"], g["c1"], g["c2"], g["obs_rot"], g["obs_t"], weight=g["weight"], jac=False)
    R = Rot.from_quat(g["rot"])
    R12 = R[g["c1"]].inv() * R[g["c2"]]
    p12 = R[g["c1"]].inv().apply(g["t"][g["c2"]] - g["t"][g["c1"]])
    dR = R12.inv() * Rot.from_quat(g["obs_rot"])
    dt = R12.inv().apply(g["obs_t"] - p12)
    ypr = dR.as_euler("ZYX", degrees=True)                                   # R = Rz(yaw) Ry(pitch) Rx(roll), CeresResidues.h:226-243
    want = g["weight"][:, None] * np.c_[dt, 4 * ypr[:, 0], 10 * ypr[:, 1], 10 * ypr[:, 2]]
    assert close(out["r"], want, 1e-10)
    assert abs(out["cost"] - 0.5 * np.sum(want ** 2)) <= 1e-10 * out["cost"]


@pytest.mark.parametrize("s", [0.0, 0.5, 0.99, 1.0])
def test_fourdof_switch_is_s_times_error_and_ignores_the_weight(s):
    g = random_edges(1, 6, 4, seed=5)
    base = pgo.fourdof_eval(0, g["rot"], g["t"], g["c1"], g["c2"], g["obs_rot"], g["obs_t"], weight=np.ones(4), jac=False)["r"]
    out = pgo.fourdof_eval(1, g["rot"], g["t"], g["c1"], g["c2"], g["obs_rot"], g["obs_t"], weight=g["weight"] * 7, sw=np.full(4, s))
    assert close(out["r"][:, :6], s * base, 1e-13) and np.allclose(out["r"][:, 6], s * (1 - s), atol=1e-15)
    assert close(out["J"][:, :6, 12], base, 1e-13) and np.allclose(out["J"][:, 6, 12], 1 - 2 * s, atol=1e-15)   # d/ds of s*e and s(1-s)


def test_qin_known_answers_and_angle_wrap():
    # yaw_i = 90 deg: R^T (t_j - t_i) with t_j - t_i = (1,0,0) is (0,-1,0)
    r, J = one(2, [90.0, 0, 0], Z3, [100.0, 0, 0], [1.0, 0, 0], [10.0, 0.0, 0.0], [0.0, -1.0, 0.0])
    assert np.abs(r).max() < 1e-15
    # wrap: 170 - (-170) - 0 = 340 -> -20 -> / 10
    r, J = one(2, [-170.0, 0, 0], Z3, [170.0, 0, 0], Z3, [0.0, 0.0, 0.0], Z3)
    assert abs(r[3] + 2.0) < 1e-15 and J[3, 0] == -0.1 and J[3, 4] == 0.1
    # the boundary is exclusive on both sides (CeresResidues.h:430-435)
    assert one(2, [0.0, 0, 0], Z3, [180.0, 0, 0], Z3, [0.0, 0, 0], Z3)[0][3] == 18.0
    assert one(2, [0.0, 0, 0], Z3, [-180.0, 0, 0], Z3, [0.0, 0, 0], Z3)[0][3] == -18.0
    # pitch and roll of the FIRST keyframe enter as constants: general case against scipy's ZYX matrix
    rng = np.random.default_rng(0)
    for _ in range(20):
        yi, p, rr, yj = rng.uniform(-180, 180), rng.uniform(-60, 60), rng.uniform(-60, 60), rng.uniform(-180, 180)
        ti, tj, to = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
        rel = rng.uniform(-90, 90)
        r, _ = one(2, [yi, 123.0, -45.0], ti, [yj, 7.0, 8.0], tj, [rel, p, rr], to)
        Ri = Rot.from_euler("ZYX", [yi, p, rr], degrees=True).as_matrix()
        d = yj - yi - rel
        d = d - 360 if d > 180 else d + 360 if d < -180 else d
        assert np.allclose(r, np.r_[Ri.T @ (tj - ti) - to, d / 10], atol=1e-12)


def test_angle_local_parameterization_and_ypr_conventions():
    assert pgo.angle_plus(170.0, 20.0) == -170.0 and pgo.angle_plus(-170.0, -20.0) == 170.0 and pgo.angle_plus(10.0, 5.0) == 15.0
    for th in (-200.0, -180.0, 0.0, 37.0, 180.0, 250.0):
        assert pgo.angle_plus_jacobian(th) == 1.0
    rng = np.random.default_rng(1)
    for _ in range(20):
        y, p, r = rng.uniform(-180, 180), rng.uniform(-89, 89), rng.uniform(-180, 180)
        R = pgo.ypr_to_R(y, p, r)
        assert np.allclose(R, Rot.from_euler("ZYX", [y, p, r], degrees=True).as_matrix(), atol=1e-14)
        M = np.eye(4); M[:3, :3] = R
        assert np.allclose(pgo.r2ypr_deg(M), [y, p, r], atol=1e-10)            # R2ypr inverts it (PoseManipUtils.cpp:143-158)


# ------------------------------------------------------------------------------------------------ Jacobians
def _plus(kind, rot, t, sw, node_delta, sw_delta):
    """Ceres' Plus on every parameter block: quaternion blocks through EigenQuaternionParameterization, yaw through
    AngleLocalParameterization, translations and switches additively."""
    rot2, t2 = rot.copy(), t.copy()
    for i, d in node_delta.items():
        if kind == 2:
            rot2[i, 0] = pgo.angle_plus(rot[i, 0], d[0]); t2[i] = t[i] + d[1:4]
        else:
            rot2[i] = pgo.quat_plus(rot[i], d[:3]); t2[i] = t[i] + d[3:6]
    return rot2, t2, (None if sw is None else sw + sw_delta)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_autodiff_jacobians_match_central_differences_through_plus(kind):
    NR, NC, _ = pgo.FOURDOF_SHAPES[kind]
    g = random_edges(kind, 12, 10, seed=11 + kind)
    out = pgo.fourdof_eval(kind, **g)
    npp = 4 if kind == 2 else 6                                              # tangent size of one pose
    h = 1e-6
    for e in range(10):
        a, b = int(g["c1"][e]), int(g["c2"][e])
        sub = {k: (v[e:e + 1] if v is not None and k not in ("rot", "t") else v) for k, v in g.items()}
        for col in range(NC):
            def shifted(sign):
                nd, sd = {}, np.zeros(1)
                if col < 2 * npp:
                    d = np.zeros(npp); d[col % npp] = sign * h; nd[a if col < npp else b] = d
                else:
                    sd = np.array([sign * h])
                rot2, t2, sw2 = _plus(kind, g["rot"], g["t"], sub["sw"], nd, sd)
                return pgo.fourdof_eval(kind, rot2, t2, sub["c1"], sub["c2"], sub["obs_rot"], sub["obs_t"], weight=sub["weight"], sw=sw2, jac=False)["r"][0]
            fd = (shifted(+1) - shifted(-1)) / (2 * h)
            if kind == 2 and col in (0, npp):
                fd = np.where(np.abs(fd) > 1e6, out["J"][e][:, col], fd)     # a +-180 wrap inside the stencil
            assert np.abs(fd - out["J"][e][:, col]).max() <= 2e-6 * max(1.0, np.abs(out["J"][e]).max()), (kind, e, col)


# ------------------------------------------------------------------------------------------------ device arithmetic on the host
@pytest.fixture(scope="module")
def hostcheck(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostcheck") / "fourdof_hostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, os.path.join(HERE, "fourdof_hostcheck.cpp")])
    return C.CDLL(so)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_device_header_arithmetic_matches_the_oracle(hostcheck, kind):
    """pgs_fourdof.cuh seeds its duals in the tangent space; the oracle differentiates the ambient parameters and
    multiplies by the Plus-Jacobians afterwards, as Ceres does.  Same numbers to rounding."""
    NR, NC, _ = pgo.FOURDOF_SHAPES[kind]
    g = random_edges(kind, 300, 2000, seed=21 + kind)
    want = pgo.fourdof_eval(kind, **g)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    arr = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    rot, t, obs_rot, obs_t, w, sw = (arr(g[k]) for k in ("rot", "t", "obs_rot", "obs_t", "weight", "sw"))
    ptr = lambda a: None if a is None else a.ctypes.data_as(dp)
    r = np.zeros((2000, NR)); J = np.zeros((2000, NR, NC))
    hostcheck.hostcheck_fourdof(C.c_int(kind), ptr(rot), ptr(t), C.c_int(2000), g["c1"].ctypes.data_as(ip), g["c2"].ctypes.data_as(ip),
                                ptr(obs_rot), ptr(obs_t), ptr(w), ptr(sw), ptr(r), ptr(J))
    assert close(r, want["r"], 1e-12) and close(J, want["J"], 1e-12)


# ------------------------------------------------------------------------------------------------ the kernel
@pytest.mark.gpu
@pytest.mark.parametrize("kind,n_edges", [(0, 5000), (1, 5000), (2, 5000), (0, 31), (1, 33), (2, 1)])
def test_device_evaluation_matches_the_oracle(kind, n_edges):
    import solve_keyframe_pose_graph_b200.capi as capi
    g = random_edges(kind, 700, n_edges, seed=31 + kind)
    want = pgo.fourdof_eval(kind, **g)
    got = capi.fourdof_evaluate(kind, **g)
    assert close(got["r"], want["r"], 1e-12) and close(got["J"], want["J"], 1e-12)
    assert abs(got["cost"] - want["cost"]) <= 1e-12 * want["cost"]
    again = capi.fourdof_evaluate(kind, **g)
    assert np.array_equal(again["r"], got["r"]) and np.array_equal(again["J"], got["J"]) and again["cost"] == got["cost"]   # reproducible
    only_r = capi.fourdof_evaluate(kind, jac=False, **g)
    assert np.array_equal(only_r["r"], got["r"]) and only_r["cost"] == got["cost"]


@pytest.mark.gpu
def test_device_evaluation_rejects_bad_input_and_accepts_an_empty_list():
    import solve_keyframe_pose_graph_b200 as pgs
    import solve_keyframe_pose_graph_b200.capi as capi
    g = random_edges(0, 10, 8, seed=41)
    bad = dict(g); bad["c2"] = g["c2"].copy(); bad["c2"][3] = 10
    with pytest.raises(pgs.PgsError, match="out of range"):
        capi.fourdof_evaluate(0, **bad)
    with pytest.raises(pgs.PgsError, match="null array"):
        capi.fourdof_evaluate(1, **g)                                         # the switched functor needs its switch array
    e = capi.fourdof_evaluate(0, g["rot"], g["t"], [], [], np.zeros((0, 4)), np.zeros((0, 3)), weight=[])
    assert e["cost"] == 0.0 and e["r"].shape == (0, 6)


# ------------------------------------------------------------------------------------------------ through the facade
def _session(dry_run):
    """A four-world session (config-4 recipe, reduced), in the facade and in the oracle's front-end restatement."""
    from oracle import frontend
    from solve_keyframe_pose_graph_b200 import facade, synth
    g = synth.generate_config(4, n_nodes=120, n_interworld=18)
    F = facade.Facade(odom_fanout=3, dry_run=dry_run); F.ingest(g)
    M = frontend.Manager(); M.ingest(g)
    R = frontend.ReferenceFrontEnd(M, odom_fanout=3)
    return g, F, R


def _same_terms(a, b):
    assert np.array_equal(a["c1"], np.asarray(b["c1"], np.int32)) and np.array_equal(a["c2"], np.asarray(b["c2"], np.int32))
    assert np.allclose(a["t"], b["t"], atol=1e-9) and np.allclose(a["obs_t"], b["obs_t"], atol=1e-9)
    if a["rot"].shape[1] == 4:                                               # quaternions up to sign
        assert np.all(np.abs(np.sum(a["rot"] * b["rot"], axis=1)) > 1 - 1e-12)
        assert np.all(np.abs(np.sum(a["obs_rot"] * b["obs_rot"], axis=1)) > 1 - 1e-12)
    else:                                                                    # degrees
        assert np.allclose(a["rot"], b["rot"], atol=1e-7) and np.allclose(a["obs_rot"], b["obs_rot"], atol=1e-7)
    for k in ("weight", "sw"):
        assert (a[k] is None) == (b[k] is None) and (a[k] is None or np.allclose(a[k], b[k], rtol=1e-9, atol=0))


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_facade_builds_the_blocks_of_the_switched_off_builds(kind):
    g, F, R = _session(dry_run=True)
    assert F.solve_once() and R.trigger(solve=False) is None
    a, b = F.alternative_terms(kind), R.alternative_terms(kind)
    n_odom, n_loop = len(R.odom), len(R.loops)
    assert len(a["c1"]) == {0: n_odom, 1: n_loop, 2: n_odom + n_loop}[kind] and len(a["c1"]) > 0
    _same_terms(a, b)
    if kind == 1:
        assert np.all(a["sw"] == 0.99)                                       # PoseGraphSLAM.cpp:353
    if kind == 2:                                                            # pitch / roll of a loop block come from paur.first = the block's SECOND pose
        own = np.array([pgo.r2ypr_deg(R.m.poses[int(i)]) for i in a["c2"][n_odom:]])
        assert np.allclose(a["obs_rot"][n_odom:, 1:], own[:, 1:], atol=1e-7)
    F.close()
