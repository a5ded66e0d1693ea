"""Property tests (hypothesis) for the oracle's functors, SURVEY §8c item 2: on arbitrary unit quaternions, translations,
weights and switch values the Jet-autodiff path (what the reference runs through ceres::AutoDiffCostFunction), the
closed-form tangent Jacobians (what the CUDA sweep implements) and a central difference through Plus agree, and the
residuals obey the symmetries the formulas imply."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pgo

finite = lambda lo, hi: st.floats(min_value=lo, max_value=hi, allow_nan=False, allow_infinity=False)
vec3 = lambda s: st.tuples(finite(-s, s), finite(-s, s), finite(-s, s)).map(np.array)


@st.composite
def unit_quat(draw):
    v = np.array(draw(st.tuples(finite(-1, 1), finite(-1, 1), finite(-1, 1), finite(-1, 1))))
    n = np.linalg.norm(v)
    return np.array([0.0, 0.0, 0.0, 1.0]) if n < 1e-3 else v / n


SET = settings(max_examples=80, deadline=None)


@SET
@given(unit_quat(), vec3(50), unit_quat(), vec3(50), unit_quat(), vec3(50), finite(0.01, 3.0))
def test_sixdof_autodiff_equals_closed_form(q1, t1, q2, t2, oq, ot, w):
    ra, Ja = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, autodiff=True)
    rc, Jc = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, autodiff=False)
    assert np.abs(ra - rc).max() <= 1e-12 * max(1.0, np.abs(ra).max())
    assert np.abs(Ja - Jc).max() <= 1e-11 * max(1.0, np.abs(Ja).max())


@SET
@given(unit_quat(), vec3(50), unit_quat(), vec3(50), unit_quat(), vec3(50), finite(-0.5, 1.5))
def test_switch_autodiff_equals_closed_form_and_factorises(q1, t1, q2, t2, oq, ot, s):
    ra, Ja = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, autodiff=True)
    rc, Jc = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, autodiff=False)
    assert np.abs(ra - rc).max() <= 1e-12 * max(1.0, np.abs(ra).max())
    assert np.abs(Ja - Jc).max() <= 1e-11 * max(1.0, np.abs(Ja).max())
    e, Je = pgo.sixdof(q1, t1, q2, t2, oq, ot, 1.0)                          # r = s [e ; 1 - s], weight ignored (CeresResidues.h:186-198)
    assert np.abs(ra[:6] - s * e).max() <= 1e-12 * max(1.0, np.abs(e).max()) and abs(ra[6] - s * (1 - s)) <= 1e-15
    assert np.abs(Ja[:6, :12] - s * Je).max() <= 1e-11 * max(1.0, np.abs(Je).max())
    assert np.abs(Ja[:6, 12] - e).max() <= 1e-12 * max(1.0, np.abs(e).max()) and abs(Ja[6, 12] - (1 - 2 * s)) <= 1e-15


@SET
@given(unit_quat(), vec3(50), unit_quat(), vec3(50), unit_quat(), vec3(50))
def test_sixdof_symmetries(q1, t1, q2, t2, oq, ot):
    r, _ = pgo.sixdof(q1, t1, q2, t2, oq, ot, 1.0)
    # q and -q are the same rotation: flipping a POSE quaternion or the OBSERVATION flips only the sign of the rotational part
    for flipped in (pgo.sixdof(-q1, t1, q2, t2, oq, ot, 1.0)[0], pgo.sixdof(q1, t1, q2, t2, -oq, ot, 1.0)[0]):
        assert np.abs(flipped[:3] - r[:3]).max() <= 1e-11 * max(1.0, np.abs(r).max())
        assert np.abs(flipped[3:] + r[3:]).max() <= 1e-12
    # a common rigid motion of both poses (gauge freedom) leaves the residual unchanged
    gq, gt = np.array([0.1, -0.2, 0.3, 0.9]), np.array([3.0, -4.0, 5.0]); gq /= np.linalg.norm(gq)
    G = pgo.pose_to_mat4(gq, gt)
    q1g, t1g = pgo.mat4_to_pose(G @ pgo.pose_to_mat4(q1, t1)); q2g, t2g = pgo.mat4_to_pose(G @ pgo.pose_to_mat4(q2, t2))
    rg, _ = pgo.sixdof(q1g, t1g, q2g, t2g, oq, ot, 1.0)
    same = np.abs(rg - r).max() <= 1e-9 * max(1.0, np.abs(r).max())
    mirrored = np.abs(rg[:3] - r[:3]).max() <= 1e-9 * max(1.0, np.abs(r).max()) and np.abs(rg[3:] + r[3:]).max() <= 1e-9
    assert same or mirrored                                                   # Quaterniond(Matrix3d) may land on the other hemisphere


@SET
@given(unit_quat(), vec3(50), unit_quat(), vec3(50), finite(0.1, 5.0))
def test_regulariser_autodiff_equals_closed_form(q, t, qf, tf, w):
    ra, Ja = pgo.node_reg(q, t, qf, tf, w, autodiff=True)
    rc, Jc = pgo.node_reg(q, t, qf, tf, w, autodiff=False)
    d = abs(float(np.dot(q, qf)))
    if d < 1e-3:                                                              # trace <= 0 branch boundary of Quaternion(Matrix3): sign may differ
        return
    assert np.abs(ra - rc).max() <= 1e-10 * max(1.0, np.abs(ra).max())
    assert np.abs(Ja - Jc).max() <= 1e-8 * max(1.0, np.abs(Ja).max())


@SET
@given(unit_quat(), vec3(0.5))
def test_plus_stays_on_the_sphere_and_its_jacobian_is_the_derivative_at_zero(q, d):
    xp = pgo.quat_plus(q, d)
    assert abs(np.linalg.norm(xp) - 1.0) <= 1e-14
    J = pgo.quat_plus_jacobian(q)
    h = 1e-6
    for c in range(3):
        e = np.zeros(3); e[c] = h
        fd = (pgo.quat_plus(q, e) - pgo.quat_plus(q, -e)) / (2 * h)
        assert np.abs(fd - J[:, c]).max() <= 1e-9
    assert np.abs(J.T @ q).max() <= 1e-15                                     # tangent columns are orthogonal to q
