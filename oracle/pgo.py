"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

ctypes front for oracle/libpgo.so (the CPU restatement of the reference's numerical
core, see pgo_core.hpp / pgo_solver.hpp).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference arm may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class Options(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("jacobi_scaling", C.c_int),
        ("use_autodiff", C.c_int),
        ("num_threads", C.c_int),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("termination", C.c_int),
        ("num_successful_steps", C.c_int),
        ("num_unsuccessful_steps", C.c_int),
        ("num_iterations", C.c_int),
        ("t_evaluate", C.c_double),
        ("t_linear", C.c_double),
        ("t_total", C.c_double),
        ("fixed_cost", C.c_double),
    ]


class Iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int),
        ("cost", C.c_double),
        ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double),
        ("gradient_norm", C.c_double),
        ("step_norm", C.c_double),
        ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double),
        ("step_is_valid", C.c_int),
        ("step_is_successful", C.c_int),
    ]


def _host_arch():
    """What -march=native means on this host (the Makefile stores the same string beside the library)."""
    try:
        out = subprocess.run(["g++", "-march=native", "-Q", "--help=target"], capture_output=True, text=True, timeout=20).stdout
        for line in out.splitlines():
            if "-march=" in line:
                return "".join(line.split())
    except (OSError, subprocess.SubprocessError):
        pass
    return None


def build(force=False):
    so = os.path.join(_HERE, "libpgo.so")
    srcs = [os.path.join(_HERE, f) for f in ("pgo_capi.cpp", "pgo_core.hpp", "pgo_solver.hpp", "pgo_fourdof.hpp")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if not stale:   # built with -march=native on another machine (the library travels with the repo snapshot)?
        try:
            built_for = open(so + ".arch").read().strip()
        except OSError:
            built_for = None
        here = _host_arch()
        stale = here is not None and built_for != here
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.pgo_create.restype = C.c_void_p
        L.pgo_evaluate.restype = C.c_double
        L.pgo_time_sweep.restype = C.c_double
        L.pgo_fourdof_eval.restype = C.c_double
        L.pgo_angle_plus.restype = C.c_double; L.pgo_angle_plus.argtypes = [C.c_double, C.c_double]
        L.pgo_angle_plus_jacobian.restype = C.c_double; L.pgo_angle_plus_jacobian.argtypes = [C.c_double]
        L.pgo_ypr_to_R.argtypes = [C.c_double, C.c_double, C.c_double, c_dp]
        for name in ("pgo_destroy", "pgo_set_nodes", "pgo_add_odom_edges", "pgo_add_loop_edges", "pgo_set_regularizers",
                     "pgo_set_switches", "pgo_get_poses", "pgo_get_switches", "pgo_evaluate", "pgo_time_sweep",
                     "pgo_linear_step", "pgo_solve"):
            getattr(L, name).argtypes = None
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def default_options(**kw):
    o = Options()
    lib().pgo_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


TERMINATION = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}


class Problem:
    """Mirror of the ceres::Problem the reference builds (parameter blocks q,t per node, one
    switch per loop edge; residual blocks SixDOFError / ...WithSwitchingConstraints /
    NodePoseRegularization)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.pgo_create())
        self.N = 0
        self.n_odom = self.n_loop = self.n_reg = 0

    def __del__(self):
        try:
            self.L.pgo_destroy(self.h)
        except Exception:
            pass

    def set_nodes(self, q, t):
        q, qp = _d(q); t, tp = _d(t)
        self.N = q.shape[0]
        self.L.pgo_set_nodes(self.h, C.c_int(self.N), qp, tp)

    def set_constant_nodes(self, first, n, constant=True):
        self.L.pgo_set_constant_nodes(self.h, C.c_int(first), C.c_int(n), C.c_int(int(constant)))

    def add_odom_edges(self, c1, c2, q, t, w):
        c1, c1p = _i(c1); c2, c2p = _i(c2); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        self.L.pgo_add_odom_edges(self.h, C.c_int(len(c1)), c1p, c2p, qp, tp, wp)
        self.n_odom += len(c1)

    def add_loop_edges(self, c1, c2, q, t, w, s_init=None):
        c1, c1p = _i(c1); c2, c2p = _i(c2); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        sp = None
        if s_init is not None:
            s_init, sp = _d(s_init)
        self.L.pgo_add_loop_edges(self.h, C.c_int(len(c1)), c1p, c2p, qp, tp, wp, sp)
        self.n_loop += len(c1)

    def set_regularizers(self, node, q, t, w):
        node, np_ = _i(node); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        self.L.pgo_set_regularizers(self.h, C.c_int(len(node)), np_, qp, tp, wp)
        self.n_reg = len(node)

    def set_switches(self, s):
        s, sp = _d(s)
        self.L.pgo_set_switches(self.h, C.c_int(len(s)), sp)

    def poses(self):
        q = np.empty((self.N, 4)); t = np.empty((self.N, 3))
        self.L.pgo_get_poses(self.h, q.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp))
        return q, t

    def switches(self):
        s = np.empty(self.n_loop)
        if self.n_loop:
            self.L.pgo_get_switches(self.h, s.ctypes.data_as(c_dp))
        return s

    def evaluate(self, autodiff=True, threads=1, jac=True):
        """Returns dict(cost, r_o[E,6], J_o[E,6,12], r_l[E,7], J_l[E,7,13], r_r[K,6], J_r[K,6,6], g_p[N,6], g_s[El])."""
        out = dict(r_o=np.zeros((self.n_odom, 6)), r_l=np.zeros((self.n_loop, 7)), r_r=np.zeros((self.n_reg, 6)))
        ptr = lambda a: a.ctypes.data_as(c_dp)
        if jac:
            out.update(J_o=np.zeros((self.n_odom, 6, 12)), J_l=np.zeros((self.n_loop, 7, 13)), J_r=np.zeros((self.n_reg, 6, 6)),
                       g_p=np.zeros((self.N, 6)), g_s=np.zeros(self.n_loop))
            cost = self.L.pgo_evaluate(self.h, C.c_int(int(autodiff)), C.c_int(threads), ptr(out["r_o"]), ptr(out["J_o"]), ptr(out["r_l"]),
                                       ptr(out["J_l"]), ptr(out["r_r"]), ptr(out["J_r"]), ptr(out["g_p"]), ptr(out["g_s"]))
        else:
            cost = self.L.pgo_evaluate(self.h, C.c_int(int(autodiff)), C.c_int(threads), ptr(out["r_o"]), None, ptr(out["r_l"]), None,
                                       ptr(out["r_r"]), None, None, None)
        out["cost"] = cost
        return out

    def time_sweep(self, autodiff=True, threads=1, reps=3, jac=True):
        return self.L.pgo_time_sweep(self.h, C.c_int(int(autodiff)), C.c_int(threads), C.c_int(reps), C.c_int(int(jac)))

    def linear_step(self, radius, options=None):
        o = options or default_options()
        dp = np.zeros((self.N, 6)); ds = np.zeros(max(self.n_loop, 1)); mcc = C.c_double(0)
        rc = self.L.pgo_linear_step(self.h, C.byref(o), C.c_double(radius), dp.ctypes.data_as(c_dp), ds.ctypes.data_as(c_dp), C.byref(mcc))
        if rc:
            raise RuntimeError("oracle linear solver failure")
        return dp, ds[: self.n_loop], mcc.value

    def solve(self, options=None):
        o = options or default_options()
        s = Summary(); cap = o.max_num_iterations + 8
        its = (Iteration * cap)()
        self.L.pgo_solve(self.h, C.byref(o), C.byref(s), its, C.c_int(cap))
        rows = [{f: getattr(its[i], f) for f, _ in Iteration._fields_} for i in range(min(s.num_iterations, cap))]
        return dict(initial_cost=s.initial_cost, final_cost=s.final_cost, termination=TERMINATION[s.termination],
                    num_successful_steps=s.num_successful_steps, num_unsuccessful_steps=s.num_unsuccessful_steps,
                    iterations=rows, t_evaluate=s.t_evaluate, t_linear=s.t_linear, t_total=s.t_total, fixed_cost=s.fixed_cost)


# ---- helper conversions (Eigen semantics) -----------------------------------------------
def mat4_to_pose(M):
    M, Mp = _d(M); q = np.empty(4); t = np.empty(3)
    lib().pgo_mat4_to_pose(Mp, q.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp))
    return q, t


def pose_to_mat4(q, t):
    q, qp = _d(q); t, tp = _d(t); M = np.empty((4, 4))
    lib().pgo_pose_to_mat4(qp, tp, M.ctypes.data_as(c_dp))
    return M


def inv4(M):
    M, Mp = _d(M); out = np.empty((4, 4))
    lib().pgo_inv4(Mp, out.ctypes.data_as(c_dp))
    return out


def r2ypr_deg(M):
    M, Mp = _d(M); y = np.empty(3)
    lib().pgo_r2ypr_deg(Mp, y.ctypes.data_as(c_dp))
    return y


def quat_plus(x, d):
    x, xp = _d(x); d, dp = _d(d); o = np.empty(4)
    lib().pgo_quat_plus(xp, dp, o.ctypes.data_as(c_dp))
    return o


def quat_plus_jacobian(x):
    x, xp = _d(x); J = np.empty((4, 3))
    lib().pgo_quat_plus_jacobian(xp, J.ctypes.data_as(c_dp))
    return J


def sixdof(q1, t1, q2, t2, oq, ot, w=1.0, autodiff=True, jac=True):
    a = [_d(v) for v in (q1, t1, q2, t2, oq, ot)]
    r = np.zeros(6); J = np.zeros((6, 12))
    lib().pgo_sixdof(C.c_int(int(autodiff)), a[0][1], a[1][1], a[2][1], a[3][1], a[4][1], a[5][1], C.c_double(w),
                     r.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp) if jac else None)
    return (r, J) if jac else r


def sixdof_switch(q1, t1, q2, t2, s, oq, ot, w=1.0, autodiff=True, jac=True):
    a = [_d(v) for v in (q1, t1, q2, t2, np.atleast_1d(s), oq, ot)]
    r = np.zeros(7); J = np.zeros((7, 13))
    lib().pgo_sixdof_switch(C.c_int(int(autodiff)), a[0][1], a[1][1], a[2][1], a[3][1], a[4][1], a[5][1], a[6][1], C.c_double(w),
                            r.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp) if jac else None)
    return (r, J) if jac else r


def node_reg(q1, t1, qf, tf, w, autodiff=True, jac=True):
    a = [_d(v) for v in (q1, t1, qf, tf)]
    r = np.zeros(6); J = np.zeros((6, 6))
    lib().pgo_node_reg(C.c_int(int(autodiff)), a[0][1], a[1][1], a[2][1], a[3][1], C.c_double(w), r.ctypes.data_as(c_dp),
                       J.ctypes.data_as(c_dp) if jac else None)
    return (r, J) if jac else r


FOURDOF_SHAPES = {0: (6, 12, 4), 1: (7, 13, 4), 2: (4, 8, 3)}   # kind -> residual rows, tangent columns, doubles per rotation record


def fourdof_eval(kind, rot, t, c1, c2, obs_rot, obs_t, weight=None, sw=None, jac=True):
    """The reference's alternative functors over an edge list (pgo_fourdof.hpp; kinds and layouts as in
    include/pgs_fourdof.h).  Returns dict(cost, r[E,NR], J[E,NR,NC])."""
    NR, NC, RW = FOURDOF_SHAPES[kind]
    rot, rotp = _d(np.asarray(rot).reshape(-1, RW)); t, tp = _d(np.asarray(t).reshape(-1, 3))
    c1, c1p = _i(c1); c2, c2p = _i(c2)
    E = len(c1)
    obs_rot, orp = _d(np.asarray(obs_rot).reshape(E, RW)); obs_t, otp = _d(np.asarray(obs_t).reshape(E, 3))
    wp = sp = None
    if weight is not None:
        weight, wp = _d(weight)
    if sw is not None:
        sw, sp = _d(sw)
    r = np.zeros((E, NR)); J = np.zeros((E, NR, NC))
    cost = lib().pgo_fourdof_eval(C.c_int(kind), C.c_int(len(rot)), rotp, tp, C.c_int(E), c1p, c2p, orp, otp, wp, sp,
                                  r.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp) if jac else None)
    return dict(cost=cost, r=r, J=J) if jac else dict(cost=cost, r=r)


def angle_plus(theta, delta):
    return lib().pgo_angle_plus(theta, delta)


def angle_plus_jacobian(theta):
    return lib().pgo_angle_plus_jacobian(theta)


def ypr_to_R(yaw, pitch, roll):
    R = np.zeros(9)
    lib().pgo_ypr_to_R(yaw, pitch, roll, R.ctypes.data_as(c_dp))
    return R.reshape(3, 3)
