"""Pins for the oracle's restatement of src/CeresResidues.h.

The reference has no numerical tests, so these known-answer tests, the independent scipy
restatement and the autodiff == closed-form == finite-difference agreement are what anchor
the oracle."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from oracle import pgo

rng = np.random.default_rng(1234)
I4 = np.array([0.0, 0.0, 0.0, 1.0])
Z3 = np.zeros(3)


def rand_quat(scale=None):
    if scale is None:
        q = rng.normal(size=4)
        return q / np.linalg.norm(q)
    return Rot.from_rotvec(rng.normal(size=3) * scale).as_quat()


def scipy_sixdof(q1, t1, q2, t2, oq, ot, w):
    """Independent restatement of CeresResidues.h:47-66 through scipy rotations (xyzw, scalar last)."""
    R1, R2, Ro = Rot.from_quat(q1), Rot.from_quat(q2), Rot.from_quat(oq)
    R12 = R1.inv() * R2
    p12 = R1.inv().apply(t2 - t1)
    dq = (R12.inv() * Ro).as_quat()  # may differ from Eigen's by a global sign
    dt = R12.inv().apply(ot - p12)
    return w * dt, w * 2 * dq[:3], dq[3]


def test_identity_gives_zero_residual():
    r, J = pgo.sixdof(I4, Z3, I4, Z3, I4, Z3, 1.0)
    assert np.all(r == 0)
    r = pgo.sixdof_switch(I4, Z3, I4, Z3, 0.99, I4, Z3, jac=False)
    assert np.allclose(r[:6], 0) and np.isclose(r[6], 0.99 * 0.01)
    r = pgo.node_reg(I4, Z3, I4, Z3, 1.1, jac=False)
    assert np.all(r == 0)


def test_pure_translation_kat():
    # w_T_c1 = I, w_T_c2 = trans(1,0,0); observation says c2 is 2 m ahead -> delta_t = (1,0,0)
    r = pgo.sixdof(I4, Z3, I4, [1, 0, 0], I4, [2, 0, 0], 1.0, jac=False)
    assert np.allclose(r, [1, 0, 0, 0, 0, 0], atol=1e-15)
    r = pgo.sixdof(I4, Z3, I4, [1, 0, 0], I4, [1, 0, 0], 0.5, jac=False)
    assert np.allclose(r, 0, atol=1e-15)


def test_pure_yaw_90_kat():
    # c2 is yawed +90deg w.r.t. c1, observation = identity: delta_q = q12^* -> 2 vec = (0,0,-sqrt 2)
    q2 = Rot.from_euler("z", 90, degrees=True).as_quat()
    r = pgo.sixdof(I4, Z3, q2, Z3, I4, Z3, 1.0, jac=False)
    assert np.allclose(r, [0, 0, 0, 0, 0, -np.sqrt(2.0)], atol=1e-15)
    # weight scales every row (CeresResidues.h:66)
    r = pgo.sixdof(I4, Z3, q2, Z3, I4, Z3, 0.25, jac=False)
    assert np.allclose(r, [0, 0, 0, 0, 0, -0.25 * np.sqrt(2.0)], atol=1e-15)


def test_observation_sign_flip():
    q1, q2, oq = rand_quat(), rand_quat(), rand_quat()
    t1, t2, ot = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
    ra = pgo.sixdof(q1, t1, q2, t2, oq, ot, 0.7, jac=False)
    rb = pgo.sixdof(q1, t1, q2, t2, -oq, ot, 0.7, jac=False)
    assert np.allclose(ra[:3], rb[:3], atol=1e-15) and np.allclose(ra[3:], -rb[3:], atol=1e-15)


@pytest.mark.parametrize("s", [0.0, 0.5, 0.99, 1.0, -0.3, 1.7])
def test_switch_values(s):
    q1, q2, oq = rand_quat(), rand_quat(), rand_quat()
    t1, t2, ot = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
    e = pgo.sixdof(q1, t1, q2, t2, oq, ot, 1.0, jac=False)
    r, J = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, w=123.0)  # weight must be ignored (CeresResidues.h:198)
    assert np.allclose(r[:6], s * e, atol=1e-14)
    assert np.isclose(r[6], s * (1 - s), atol=1e-15)
    assert np.allclose(J[:6, 12], e, atol=1e-14) and np.isclose(J[6, 12], 1 - 2 * s, atol=1e-14)
    assert np.allclose(J[6, :12], 0)


def test_against_scipy_restatement():
    for _ in range(200):
        q1, q2, oq = rand_quat(), rand_quat(), rand_quat()
        t1, t2, ot = rng.normal(size=3) * 10, rng.normal(size=3) * 10, rng.normal(size=3)
        w = rng.uniform(0.1, 2)
        r = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, jac=False)
        dt, dqv, dqw = scipy_sixdof(q1, t1, q2, t2, oq, ot, w)
        assert np.allclose(r[:3], dt, atol=1e-12)
        # scipy canonicalises nothing but its product sign can differ from Hamilton q12* (x) qo: compare up to sign
        assert np.allclose(r[3:], dqv, atol=1e-12) or np.allclose(r[3:], -dqv, atol=1e-12)


def fd_tangent(fun, q_list, t_list, extra, m, h=1e-6):
    """Central differences through ceres::EigenQuaternionParameterization::Plus."""
    cols = []
    for b in range(len(q_list)):
        for k in range(3):
            d = np.zeros(3); d[k] = h
            qp = [q.copy() for q in q_list]; qm = [q.copy() for q in q_list]
            qp[b] = pgo.quat_plus(q_list[b], d); qm[b] = pgo.quat_plus(q_list[b], -d)
            cols.append((fun(qp, t_list, extra) - fun(qm, t_list, extra)) / (2 * h))
        for k in range(3):
            tp = [t.copy() for t in t_list]; tm = [t.copy() for t in t_list]
            tp[b][k] += h; tm[b][k] -= h
            cols.append((fun(q_list, tp, extra) - fun(q_list, tm, extra)) / (2 * h))
    return np.stack(cols, axis=1)


def test_sixdof_jacobian_autodiff_closed_fd():
    worst = 0
    for it in range(100):
        sc = None if it % 2 else 0.3
        q1, q2, oq = rand_quat(sc), rand_quat(sc), rand_quat(sc)
        t1, t2, ot = rng.normal(size=3) * 5, rng.normal(size=3) * 5, rng.normal(size=3)
        w = rng.uniform(0.1, 2)
        ra, Ja = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, autodiff=True)
        rc, Jc = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, autodiff=False)
        assert np.allclose(ra, rc, rtol=0, atol=1e-12)
        assert np.allclose(Ja, Jc, rtol=0, atol=1e-11)
        f = lambda qs, ts, _: pgo.sixdof(qs[0], ts[0], qs[1], ts[1], oq, ot, w, jac=False)
        Jf = fd_tangent(f, [q1, q2], [t1, t2], None, 6)
        worst = max(worst, np.abs(Jf - Ja).max())
    assert worst < 2e-8, worst


def test_switch_jacobian_autodiff_closed_fd():
    for it in range(60):
        q1, q2, oq = rand_quat(), rand_quat(), rand_quat()
        t1, t2, ot = rng.normal(size=3) * 5, rng.normal(size=3) * 5, rng.normal(size=3)
        s = rng.uniform(-0.2, 1.2)
        ra, Ja = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, autodiff=True)
        rc, Jc = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, autodiff=False)
        assert np.allclose(ra, rc, rtol=0, atol=1e-12) and np.allclose(Ja, Jc, rtol=0, atol=1e-11)
        f = lambda qs, ts, ss: pgo.sixdof_switch(qs[0], ts[0], qs[1], ts[1], ss, oq, ot, jac=False)
        Jf = fd_tangent(f, [q1, q2], [t1, t2], s, 7)
        assert np.abs(Jf - Ja[:, :12]).max() < 2e-8
        h = 1e-6
        js = (f([q1, q2], [t1, t2], s + h) - f([q1, q2], [t1, t2], s - h)) / (2 * h)
        assert np.abs(js - Ja[:, 12]).max() < 1e-8


def test_regulariser_autodiff_closed_fd_and_sign_branches():
    for it in range(120):
        qf = rand_quat()
        # half the cases close to the anchor (trace>0 branch), half far away (other Shepperd branches)
        q = pgo.quat_plus(qf, rng.normal(size=3) * 0.05) if it % 2 == 0 else rand_quat()
        tf, t = rng.normal(size=3) * 5, rng.normal(size=3) * 5
        w = rng.uniform(1.1, 6)
        ra, Ja = pgo.node_reg(q, t, qf, tf, w, autodiff=True)
        rc, Jc = pgo.node_reg(q, t, qf, tf, w, autodiff=False)
        assert np.allclose(ra, rc, rtol=0, atol=1e-11), (it, ra, rc)
        assert np.allclose(Ja, Jc, rtol=0, atol=1e-10)
        f = lambda qs, ts, _: pgo.node_reg(qs[0], ts[0], qf, tf, w, jac=False)
        Jf = fd_tangent(f, [q], [t], None, 6)
        assert np.abs(Jf - Ja).max() < 5e-8
    r = pgo.node_reg(qf, tf, qf, tf, 3.0, jac=False)
    assert np.allclose(r, 0, atol=1e-13)   # w * |t| * a few ulp: translations here are ~5 m, the weight 3


def test_plus_and_plus_jacobian():
    for _ in range(50):
        x = rand_quat()
        J = pgo.quat_plus_jacobian(x)
        h = 1e-6
        Jf = np.stack([(pgo.quat_plus(x, h * np.eye(3)[k]) - pgo.quat_plus(x, -h * np.eye(3)[k])) / (2 * h) for k in range(3)], axis=1)
        assert np.abs(J - Jf).max() < 1e-9
        d = rng.normal(size=3) * 0.4
        xp = pgo.quat_plus(x, d)
        assert abs(np.linalg.norm(xp) - 1) < 1e-14
        # left multiplication by a rotation of angle 2|d| about d
        want = (Rot.from_rotvec(2 * d) * Rot.from_quat(x)).as_quat()
        assert np.allclose(xp, want, atol=1e-13) or np.allclose(xp, -want, atol=1e-13)
    assert np.array_equal(pgo.quat_plus(x, np.zeros(3)), x)


def test_eigen_conversions():
    # Quaternion(Matrix3) Shepperd branches + toRotationMatrix round trip, w>0 in the trace>0 branch
    for _ in range(200):
        q = rand_quat(); t = rng.normal(size=3)
        M = pgo.pose_to_mat4(q, t)
        assert np.allclose(M[:3, :3], Rot.from_quat(q).as_matrix(), atol=1e-14)
        q2, t2 = pgo.mat4_to_pose(M)
        assert np.allclose(t2, t)
        assert np.allclose(q2, q, atol=1e-13) or np.allclose(q2, -q, atol=1e-13)
        if np.trace(M[:3, :3]) > 0:
            assert q2[3] > 0
        assert np.allclose(pgo.inv4(M) @ M, np.eye(4), atol=1e-13)
    # R2ypr is in DEGREES (PoseManipUtils.cpp:157)
    M = np.eye(4); M[:3, :3] = Rot.from_euler("ZYX", [30, 10, -5], degrees=True).as_matrix()
    assert np.allclose(pgo.r2ypr_deg(M), [30, 10, -5], atol=1e-12)
