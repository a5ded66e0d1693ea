"""The oracle's functors against THE REFERENCE'S OWN FUNCTOR SOURCE.

oracle/_ref/libref_functors.so is /root/reference/src/CeresResidues.h compiled unmodified, from where it lies, over
oracle/shim/ — a stand-in for the part of the Eigen / Ceres API that file uses (neither library exists in this
container).  What executes is the reference's text: which quaternion is conjugated, what is subtracted from what, the
residual layout and its scaling, the switch penalty, the yaw/pitch/roll extraction, the angle wrap.  The shim supplies the
Eigen primitives underneath from their documented semantics, so this pins the oracle's TRANSCRIPTION of the functors, not
Eigen's or Ceres' arithmetic (oracle/shim/mini_eigen.hpp).  Jacobians: the reference's templates instantiated with Jets,
times the Plus-Jacobian, on both sides.  Built only where the reference tree exists; the tests skip without the library."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pgo

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_functors.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_functors.so not built (needs /root/reference; make -C oracle)")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(REF_SO)
    L.ref_sixdof.argtypes = [dp, dp, dp, dp, dp, C.c_double, dp, dp]
    L.ref_fourdof.argtypes = [dp, dp, dp, dp, dp, C.c_double, dp, dp]
    L.ref_sixdof_switch.argtypes = [dp, dp, dp, dp, dp, dp, C.c_double, dp, dp]
    L.ref_fourdof_switch.argtypes = [dp, dp, dp, dp, dp, dp, C.c_double, dp, dp]
    L.ref_node_reg.argtypes = [dp, dp, dp, C.c_double, dp, dp]
    L.ref_qin.argtypes = [C.c_double, dp, C.c_double, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp]
    L.ref_normalize_angle.restype = C.c_double; L.ref_normalize_angle.argtypes = [C.c_double]
    L.ref_angle_plus.restype = C.c_double; L.ref_angle_plus.argtypes = [C.c_double, C.c_double]
    L.ref_ypr_to_R.argtypes = [C.c_double, C.c_double, C.c_double, dp]
    L.ref_r2ypr.argtypes = [dp, dp]
    return L


def ptr(a):
    return a.ctypes.data_as(dp)


def arr(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def rq(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def close(a, b, tol=1e-13):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * max(1.0, np.abs(b).max())


def case(rng, small_error):
    """Two poses and an observation c1_T_c2 = true relative pose x error (small: near the solution; large: anywhere)."""
    q1, q2, t1, t2 = rq(rng), rq(rng), rng.normal(size=3) * 10, rng.normal(size=3) * 10
    T1, T2 = pgo.pose_to_mat4(q1, t1), pgo.pose_to_mat4(q2, t2)
    E = pgo.pose_to_mat4(pgo.quat_plus(np.array([0, 0, 0, 1.0]), rng.normal(size=3) * (0.05 if small_error else 1.0)),
                         rng.normal(size=3) * (0.1 if small_error else 5.0))
    return arr(q1), arr(t1), arr(q2), arr(t2), arr(pgo.inv4(T1) @ T2 @ E)


@pytest.mark.parametrize("small_error", [True, False])
def test_sixdof_and_its_switching_variant_equal_the_reference_source(ref, small_error):
    rng = np.random.default_rng(1 + small_error)
    for _ in range(200):
        q1, t1, q2, t2, obs = case(rng, small_error)
        w, s = rng.uniform(0.1, 2.0), rng.uniform(-0.3, 1.3)
        oq, ot = pgo.mat4_to_pose(obs)                                   # what the reference's constructor stores (CeresResidues.h:22-28)
        r = np.zeros(6); J = np.zeros((6, 12))
        ref.ref_sixdof(ptr(q1), ptr(t1), ptr(q2), ptr(t2), ptr(obs), w, ptr(r), ptr(J))
        ro, Jo = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, autodiff=True)
        assert close(ro, r) and close(Jo, J)
        rc, Jc = pgo.sixdof(q1, t1, q2, t2, oq, ot, w, autodiff=False)   # the closed form the CUDA sweep implements
        assert close(rc, r, 1e-12) and close(Jc, J, 1e-11)
        r7 = np.zeros(7); J7 = np.zeros((7, 13)); sv = arr([s])
        ref.ref_sixdof_switch(ptr(q1), ptr(t1), ptr(q2), ptr(t2), ptr(sv), ptr(obs), w, ptr(r7), ptr(J7))
        ro, Jo = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, w, autodiff=True)
        assert close(ro, r7) and close(Jo, J7)
        rc, Jc = pgo.sixdof_switch(q1, t1, q2, t2, s, oq, ot, w, autodiff=False)
        assert close(rc, r7, 1e-12) and close(Jc, J7, 1e-11)


def test_node_regulariser_equals_the_reference_source(ref):
    rng = np.random.default_rng(3)
    for k in range(200):
        q, t, qf, tf = arr(rq(rng)), arr(rng.normal(size=3) * 10), rq(rng), rng.normal(size=3) * 10
        if k % 2:                                                        # near the anchor, where the solver uses it
            qf = pgo.quat_plus(q, rng.normal(size=3) * 0.05); tf = t + rng.normal(size=3) * 0.1
        w = rng.uniform(1.1, 4.0)
        anchor = arr(pgo.pose_to_mat4(qf, tf))
        r = np.zeros(6); J = np.zeros((6, 6))
        ref.ref_node_reg(ptr(q), ptr(t), ptr(anchor), w, ptr(r), ptr(J))
        ro, Jo = pgo.node_reg(q, t, qf, tf, w, autodiff=True)
        assert close(ro, r, 1e-12) and close(Jo, J, 1e-11)
        if abs(float(np.dot(q, qf))) > 0.3:                              # the closed form's sign rule: away from the trace <= 0 branch
            rc, Jc = pgo.node_reg(q, t, qf, tf, w, autodiff=False)
            assert close(rc, r, 1e-11) and close(Jc, J, 1e-9)


def test_switched_off_functors_equal_the_reference_source(ref):
    rng = np.random.default_rng(4)
    for _ in range(200):
        q1, t1, q2, t2, obs = case(rng, True)
        w, s = rng.uniform(0.1, 2.0), rng.uniform(-0.3, 1.3)
        oq, ot = pgo.mat4_to_pose(obs)
        r = np.zeros(6); J = np.zeros((6, 12))
        ref.ref_fourdof(ptr(q1), ptr(t1), ptr(q2), ptr(t2), ptr(obs), w, ptr(r), ptr(J))
        o = pgo.fourdof_eval(0, np.array([q1, q2]), np.array([t1, t2]), [0], [1], [oq], [ot], weight=[w])
        assert close(o["r"][0], r, 1e-12) and close(o["J"][0], J, 1e-12)
        r7 = np.zeros(7); J7 = np.zeros((7, 13)); sv = arr([s])
        ref.ref_fourdof_switch(ptr(q1), ptr(t1), ptr(q2), ptr(t2), ptr(sv), ptr(obs), w, ptr(r7), ptr(J7))
        o = pgo.fourdof_eval(1, np.array([q1, q2]), np.array([t1, t2]), [0], [1], [oq], [ot], weight=[w], sw=[s])
        assert close(o["r"][0], r7, 1e-12) and close(o["J"][0], J7, 1e-12)
        yi, yj, p, rr, rel = rng.uniform(-180, 180), rng.uniform(-180, 180), rng.uniform(-60, 60), rng.uniform(-60, 60), rng.uniform(-400, 400)
        r4 = np.zeros(4); J4 = np.zeros((4, 8)); tobs = arr(rng.normal(size=3))
        ref.ref_qin(yi, ptr(t1), yj, ptr(t2), ptr(tobs), rel, p, rr, ptr(r4), ptr(J4))
        o = pgo.fourdof_eval(2, np.array([[yi, 0, 0], [yj, 0, 0]]), np.array([t1, t2]), [0], [1], [[rel, p, rr]], [tobs])
        assert close(o["r"][0], r4, 1e-13) and close(o["J"][0], J4, 1e-13)


def test_angle_helpers_equal_the_reference_source(ref):
    rng = np.random.default_rng(5)
    for a in list(rng.uniform(-720, 720, size=200)) + [180.0, -180.0, 180.0000001, -180.0000001, 0.0, 360.0, 540.0]:
        want = a - 360 if a > 180 else a + 360 if a < -180 else a
        assert ref.ref_normalize_angle(a) == want
        assert ref.ref_angle_plus(a / 2, a / 2) == pgo.angle_plus(a / 2, a / 2)
    R9 = np.zeros(9); ypr = np.zeros(3)
    for _ in range(100):
        y, p, r = rng.uniform(-180, 180), rng.uniform(-89, 89), rng.uniform(-180, 180)
        ref.ref_ypr_to_R(y, p, r, ptr(R9))
        assert np.array_equal(R9.reshape(3, 3), pgo.ypr_to_R(y, p, r))
        ref.ref_r2ypr(ptr(R9), ptr(ypr))
        M = np.eye(4); M[:3, :3] = R9.reshape(3, 3)
        assert np.allclose(ypr, pgo.r2ypr_deg(M), rtol=0, atol=1e-12) and np.allclose(ypr, [y, p, r], atol=1e-9)


# ------------------------------------------------------------------------------------------------ PoseManipUtils (real source)
@pytest.fixture(scope="module")
def pmu(ref):
    ref.ref_pmu_raw_xyzw_to_eigenmat.argtypes = [dp, dp, dp]
    ref.ref_pmu_eigenmat_to_raw_xyzw.argtypes = [dp, dp, dp]
    ref.ref_pmu_rawyprt_to_eigenmat.argtypes = [dp, dp, dp]
    ref.ref_pmu_eigenmat_to_rawyprt.argtypes = [dp, dp, dp]
    ref.ref_pmu_prettyprint.argtypes = [dp, C.c_char_p, C.c_int]
    ref.ref_pmu_string_to_eigenmat.argtypes = [C.c_char_p, dp]
    return ref


def test_pose_conversions_equal_the_reference_pose_manip_utils(pmu):
    """src/utils/PoseManipUtils.cpp compiled from where it lies (over the same shim): (q xyzw, t) <-> 4x4 and
    (yaw, pitch, roll degrees, t) <-> 4x4 against the oracle's helpers, both hemispheres and all branches of
    Quaterniond(Matrix3d)."""
    rng = np.random.default_rng(6)
    M = np.zeros(16); q2 = np.zeros(4); t2 = np.zeros(3); ypr = np.zeros(3)
    for k in range(300):
        q, t = arr(rq(rng)), arr(rng.normal(size=3) * 10)
        if k % 3 == 0:                                                   # rotations by ~pi: the trace <= 0 branches
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = np.pi - rng.uniform(0, 0.05)
            q = arr(np.r_[np.sin(ang / 2) * ax, np.cos(ang / 2)])
        pmu.ref_pmu_raw_xyzw_to_eigenmat(ptr(q), ptr(t), ptr(M))
        assert np.array_equal(M.reshape(4, 4), pgo.pose_to_mat4(q, t))
        pmu.ref_pmu_eigenmat_to_raw_xyzw(ptr(M), ptr(q2), ptr(t2))
        qo, to = pgo.mat4_to_pose(M.reshape(4, 4))
        assert np.array_equal(q2, qo) and np.array_equal(t2, to) and abs(abs(float(np.dot(q2, q))) - 1) < 1e-12
        pmu.ref_pmu_eigenmat_to_rawyprt(ptr(M), ptr(ypr), ptr(t2))
        assert np.allclose(ypr, pgo.r2ypr_deg(M.reshape(4, 4)), rtol=0, atol=1e-12)
        M2 = np.zeros(16)
        pmu.ref_pmu_rawyprt_to_eigenmat(ptr(ypr), ptr(t), ptr(M2))      # ypr2R = Rz Ry Rx as three products (:162-187)
        assert np.allclose(M2.reshape(4, 4)[:3, :3], pgo.ypr_to_R(*ypr), rtol=0, atol=1e-15)
        assert np.allclose(M2.reshape(4, 4), M.reshape(4, 4), rtol=0, atol=1e-12)


def test_matrix_text_formats_equal_the_reference_pose_manip_utils(pmu):
    """The `data_pretty` printer and the `a,b,c,d;...` parser of the reference, byte for byte / value for value, against
    the product's (csrc/host/GraphIO.cpp through the C-ABI) — including values that print as -0.000 and 10.000."""
    from solve_keyframe_pose_graph_b200 import facade
    rng = np.random.default_rng(7)
    buf = C.create_string_buffer(256); M2 = np.zeros(16)
    for k in range(300):
        q, t = rq(rng), rng.normal(size=3) * 10.0 ** rng.integers(-5, 3)
        if k % 4 == 0:
            t[rng.integers(0, 3)] = -1e-5                                # "-0.000"
        if k % 5 == 0:
            q = np.array([0, 0, 0, 1.0]) if k % 10 else pgo.quat_plus(np.array([0, 0, 0, 1.0]), [0, 0, 1e-7])
        M = arr(pgo.pose_to_mat4(q, t))
        assert pmu.ref_pmu_prettyprint(ptr(M), buf, 256) > 0
        assert facade.io_prettyprint(M) == buf.value.decode()
        for layout in (False, True):                                     # log_posegraph.json and solved_posegraph.json spellings
            s = facade.io_mat_to_string(M, solved_layout=layout)
            ours = facade.io_string_to_mat(s)                            # 16 significant digits (Eigen FullPrecision): close, not lossless
            assert np.allclose(ours, M.reshape(4, 4), rtol=1e-15, atol=1e-300)
            if not layout:                                               # the reference's parser reads the ';' layout (:272-295)
                assert pmu.ref_pmu_string_to_eigenmat(s.encode(), ptr(M2)) == 1 and np.array_equal(M2.reshape(4, 4), ours)
    # the strings quoted in the reference's own source parse the same way in both
    import json
    sample = json.load(open(os.path.join(HERE, "golden", "reference_solved_posegraph_sample.json")))
    for node in sample["SolvedPoseGraph"]:
        T = facade.io_string_to_mat(node["w_T_c"]["data"])
        assert pmu.ref_pmu_prettyprint(ptr(arr(T)), buf, 256) > 0 and buf.value.decode() == node["w_T_c"]["data_pretty"]
