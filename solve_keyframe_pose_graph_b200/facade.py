"""ctypes view of include/pgs_facade.h: the ROS-free NodeDataManager + PoseGraphSLAM pair.

`Facade` mirrors how reference src/keyframe_pose_graph_slam_node.cpp drives the two classes: feed
keyframe poses / loop edges / kidnap signals into the manager, run one wake-up of
`reinit_ceres_problem_onnewloopedge_optimize6DOF()`, read `getNodePose` back."""
import ctypes as C

import numpy as np

from .capi import Iteration, Options, PgsError, Summary, TERMINATION, c_dp, c_ip, lib


class FacadeOptions(C.Structure):
    _fields_ = [("odom_fanout", C.c_int32), ("derive_odometry", C.c_int32), ("dry_run", C.c_int32), ("solver", Options)]


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


class Facade:
    def __init__(self, odom_fanout=5, derive_odometry=True, dry_run=False, **solver_opts):
        self.L = lib()
        self.L.pgs_facade_last_error.restype = C.c_char_p
        self.L.pgs_facade_last_error.argtypes = [C.c_void_p]
        o = FacadeOptions()
        self.L.pgs_facade_default_options(C.byref(o))
        o.odom_fanout = odom_fanout; o.derive_odometry = int(derive_odometry); o.dry_run = int(dry_run)
        for k, v in solver_opts.items():
            if not hasattr(o.solver, k):
                raise AttributeError(k)
            setattr(o.solver, k, v)
        self.opt = o
        self.h = C.c_void_p()
        if self.L.pgs_facade_create(C.byref(o), C.byref(self.h)) != 0:
            raise PgsError("pgs_facade_create failed")
        self.n_loop = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.pgs_facade_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise PgsError(f"facade error {rc}: {self.L.pgs_facade_last_error(self.h).decode(errors="replace")}")
        return rc

    # ---- ingest
    def add_nodes(self, stamps, q, t):
        stamps = np.ascontiguousarray(stamps, dtype=np.int64); q, qp = _d(q); t, tp = _d(t)
        self._ck(self.L.pgs_facade_add_nodes(self.h, C.c_int32(len(stamps)), stamps.ctypes.data_as(C.c_void_p), qp, tp))

    def add_loop_edges(self, a, b, q, t, w):
        a, ap = _i(a); b, bp = _i(b); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        self._ck(self.L.pgs_facade_add_loop_edges(self.h, C.c_int32(len(a)), ap, bp, qp, tp, wp)); self.n_loop += len(a)

    def add_loop_edge_stamped(self, sa, sb, q, t, w=1.0):
        q, qp = _d(q); t, tp = _d(t)
        r = self._ck(self.L.pgs_facade_add_loop_edge_stamped(self.h, C.c_int64(sa), C.c_int64(sb), qp, tp, C.c_double(w)))
        self.n_loop += r
        return bool(r)

    def kidnap_indicator(self, stamp, kidnapped):
        self._ck(self.L.pgs_facade_kidnap_indicator(self.h, C.c_int64(int(stamp)), C.c_int32(int(kidnapped))))

    def add_odometry_edge(self, a, b, q, t, w):
        q, qp = _d(q); t, tp = _d(t)
        self._ck(self.L.pgs_facade_add_odometry_edge(self.h, C.c_int32(a), C.c_int32(b), qp, tp, C.c_double(w)))

    # ---- ROS-message entry points (reference callback names, csrc/host/RosShim.h)
    def camera_pose_callback(self, stamp_ns, position, orientation_xyzw, covariance=None):
        p, pp = _d(position); q, qp = _d(orientation_xyzw)
        cp = None
        if covariance is not None:
            cov, cp = _d(np.asarray(covariance).reshape(36))
        self._ck(self.L.pgs_facade_camera_pose_callback(self.h, C.c_uint32(int(stamp_ns) // 10**9), C.c_uint32(int(stamp_ns) % 10**9), pp, qp, cp))

    def loopclosure_pose_callback(self, stamp0_ns, stamp1_ns, position, orientation_xyzw, weight=1.0, description=""):
        p, pp = _d(position); q, qp = _d(orientation_xyzw)
        r = self._ck(self.L.pgs_facade_loopclosure_pose_callback(self.h, C.c_uint32(int(stamp0_ns) // 10**9), C.c_uint32(int(stamp0_ns) % 10**9),
                                                                  C.c_uint32(int(stamp1_ns) // 10**9), C.c_uint32(int(stamp1_ns) % 10**9), pp, qp, C.c_float(weight), description.encode()))
        self.n_loop += r
        return bool(r)

    def rcvd_kidnap_indicator_callback(self, stamp_ns, frame_id):
        return self.L.pgs_facade_rcvd_kidnap_indicator_callback(self.h, C.c_uint32(int(stamp_ns) // 10**9), C.c_uint32(int(stamp_ns) % 10**9), frame_id.encode()) == 0

    def ingest(self, g):
        """Feed a generated graph (synth.generate) in time order: nodes, kidnap signals, then loop edges."""
        ev = sorted([(int(s), 1) for s in g["k0"]] + [(int(s), 0) for s in g["k1"]])
        pos = 0
        for stamp, kid in ev:
            # nodes with stamp <= kidnap-start belong before the signal; un-kidnap precedes the next world's first node
            cut = int(np.searchsorted(g["stamps"], stamp, side="right"))
            if cut > pos:
                self.add_nodes(g["stamps"][pos:cut], g["q"][pos:cut], g["t"][pos:cut]); pos = cut
            self.kidnap_indicator(stamp, kid)
        if pos < g["N"]:
            self.add_nodes(g["stamps"][pos:], g["q"][pos:], g["t"][pos:])
        if len(g["la"]):
            self.add_loop_edges(g["la"], g["lb"], g["lq"], g["lt"], g["lw"])

    # ---- solve
    def load_state(self):
        self._ck(self.L.pgs_facade_load_state(self.h))

    def solve_once(self, force=False):
        return self._ck(self.L.pgs_facade_solve_once(self.h, C.c_int32(int(force)))) == 1

    def thread_start(self, rate_hz=0.0):
        self._ck(self.L.pgs_facade_thread_start(self.h, C.c_double(rate_hz)))

    def thread_stop(self):
        return self._ck(self.L.pgs_facade_thread_stop(self.h))

    def status(self):
        return self.L.pgs_facade_status(self.h)

    # ---- results
    def n_nodes(self):
        return self.L.pgs_facade_n_nodes(self.h)

    def solved_until(self):
        return self.L.pgs_facade_solved_until(self.h)

    def poses(self):
        n = self.n_nodes(); q = np.zeros((max(n, 1), 4)); t = np.zeros((max(n, 1), 3))
        m = self._ck(self.L.pgs_facade_get_poses(self.h, C.c_int32(n), q.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp)))
        return q[:m], t[:m]

    def switches(self):
        s = np.zeros(max(self.n_loop, 1))
        self._ck(self.L.pgs_facade_get_switches(self.h, C.c_int32(self.n_loop), s.ctypes.data_as(c_dp)))
        return s[: self.n_loop]

    def summary(self):
        s = Summary(); cap = self.opt.solver.max_num_iterations + 8; its = (Iteration * cap)()
        self._ck(self.L.pgs_facade_get_summary(self.h, C.byref(s), its, C.c_int32(cap)))
        d = {f: getattr(s, f) for f, _ in Summary._fields_}
        d["termination"] = TERMINATION.get(s.termination, "?")
        d["iterations"] = [{f: getattr(its[i], f) for f, _ in Iteration._fields_} for i in range(min(s.num_iterations, cap))]
        return d

    # ---- Composer (reference src/Composer.cpp:10-292)
    def n_keyframes(self):
        return self.L.pgs_facade_n_keyframes(self.h)

    def compose(self):
        """One pass of Composer::pose_assember_thread on the device -> (T [n,4,4] assembled poses, world id [n])."""
        n = self.n_keyframes()
        T = np.zeros((max(n, 1), 4, 4)); w = np.zeros(max(n, 1), np.int32)
        r = self._ck(self.L.pgs_facade_compose(self.h, C.c_int32(n), T.ctypes.data_as(c_dp), w.ctypes.data_as(c_ip)))
        return T[:r], w[:r]

    def last_known_camerapose(self):
        T = np.zeros((4, 4)); st = C.c_int64(0)
        r = self.L.pgs_facade_last_known_camerapose(self.h, T.ctypes.data_as(c_dp), C.byref(st))
        return r, T, st.value

    def compose_timing(self):
        a = C.c_double(0); b = C.c_double(0)
        self._ck(self.L.pgs_facade_compose_timing(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- on-disk formats (reference log_posegraph.json / log_optimized_poses.json / solved_posegraph.json)
    def save_json(self, directory):
        return self._ck(self.L.pgs_facade_save_json(self.h, str(directory).encode()))

    def load_worlds_state(self, solved_posegraph_json):
        self._ck(self.L.pgs_facade_load_worlds_state(self.h, str(solved_posegraph_json).encode()))

    def save_state_to_disk(self, directory):
        """Composer::saveStateToDisk of the reference: end the current world at the last keyframe, write <directory>/solved_posegraph.json."""
        self._ck(self.L.pgs_facade_save_state_to_disk(self.h, str(directory).encode()))

    def load_state_from_disk(self, directory):
        """Composer::loadStateFromDisk of the reference: restore a session from <directory>/solved_posegraph.json."""
        self._ck(self.L.pgs_facade_load_state_from_disk(self.h, str(directory).encode()))

    def load_posegraph_json(self, directory):
        self._ck(self.L.pgs_facade_load_posegraph_json(self.h, str(directory).encode()))

    # ---- introspection
    def odom_terms(self):
        n = self._ck(self.L.pgs_facade_n_odom_terms(self.h))
        u = np.zeros(n, np.int32); v = np.zeros(n, np.int32); q = np.zeros((n, 4)); t = np.zeros((n, 3)); w = np.zeros(n)
        m = self._ck(self.L.pgs_facade_get_odom_terms(self.h, C.c_int32(n), u.ctypes.data_as(c_ip), v.ctypes.data_as(c_ip), q.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp), w.ctypes.data_as(c_dp)))
        return dict(u=u[:m], umf=v[:m], q=q[:m], t=t[:m], w=w[:m])

    def reg_terms(self):
        n = self._ck(self.L.pgs_facade_n_reg_terms(self.h))
        node = np.zeros(n, np.int32); q = np.zeros((n, 4)); t = np.zeros((n, 3)); w = np.zeros(n)
        m = self._ck(self.L.pgs_facade_get_reg_terms(self.h, C.c_int32(n), node.ctypes.data_as(c_ip), q.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp), w.ctypes.data_as(c_dp)))
        return dict(node=node[:m], q=q[:m], t=t[:m], w=w[:m])

    def alternative_terms(self, kind):
        """Blocks of the reference's switched-off builds for this session (PoseGraphSLAM::alternative_terms), as the
        keyword arguments of capi.fourdof_evaluate / oracle.pgo.fourdof_eval."""
        nn = C.c_int32(0); ne = C.c_int32(0)
        self._ck(self.L.pgs_facade_alternative_terms_size(self.h, C.c_int32(kind), C.byref(nn), C.byref(ne)))
        n, m = nn.value, ne.value
        rw = 3 if kind == 2 else 4
        rot = np.zeros((n, rw)); t = np.zeros((n, 3)); c1 = np.zeros(m, np.int32); c2 = np.zeros(m, np.int32)
        obs_rot = np.zeros((m, rw)); obs_t = np.zeros((m, 3)); weight = np.zeros(m); sw = np.zeros(m)
        p = lambda a: a.ctypes.data_as(c_dp)
        self._ck(self.L.pgs_facade_get_alternative_terms(self.h, C.c_int32(kind), C.c_int32(n), C.c_int32(m), p(rot), p(t), c1.ctypes.data_as(c_ip),
                                                         c2.ctypes.data_as(c_ip), p(obs_rot), p(obs_t), p(weight), p(sw)))
        return dict(rot=rot, t=t, c1=c1, c2=c2, obs_rot=obs_rot, obs_t=obs_t, weight=weight if kind in (0, 1) else None, sw=sw if kind == 1 else None)

    def evaluate_alternative(self, kind, jac=True):
        """Residuals, tangent Jacobians and cost of those blocks at the current optimisation variables, on the device."""
        nn = C.c_int32(0); ne = C.c_int32(0)
        self._ck(self.L.pgs_facade_alternative_terms_size(self.h, C.c_int32(kind), C.byref(nn), C.byref(ne)))
        nr, nc = {0: (6, 12), 1: (7, 13), 2: (4, 8)}[kind]
        r = np.zeros((ne.value, nr)); J = np.zeros((ne.value, nr, nc)) if jac else None
        cost = C.c_double(0)
        self._ck(self.L.pgs_facade_evaluate_alternative(self.h, C.c_int32(kind), C.c_int32(ne.value), r.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp) if jac else None, C.byref(cost)))
        return dict(cost=cost.value, r=r, J=J)

    def which_world(self, stamp):
        return self.L.pgs_facade_which_world(self.h, C.c_int64(int(stamp)))

    def n_worlds(self):
        return self.L.pgs_facade_n_worlds(self.h)

    def world_setid(self, w):
        return self.L.pgs_facade_world_setid(self.h, C.c_int32(w))

    def world_start(self, w):
        return self.L.pgs_facade_world_start(self.h, C.c_int32(w))

    def world_end(self, w):
        return self.L.pgs_facade_world_end(self.h, C.c_int32(w))

    def pose_between_worlds(self, m, n):
        M = np.zeros((4, 4))
        ok = self.L.pgs_facade_pose_between_worlds(self.h, C.c_int32(m), C.c_int32(n), M.ctypes.data_as(c_dp))
        return M if ok else None


def io_prettyprint(T):
    T = np.ascontiguousarray(T, dtype=np.float64); buf = C.create_string_buffer(256)
    lib().pgs_io_prettyprint(T.ctypes.data_as(c_dp), buf, C.c_int32(256))
    return buf.value.decode(errors="replace")


def io_mat_to_string(T, solved_layout=False):
    T = np.ascontiguousarray(T, dtype=np.float64); buf = C.create_string_buffer(1024)
    lib().pgs_io_mat_to_string(T.ctypes.data_as(c_dp), C.c_int32(int(solved_layout)), buf, C.c_int32(1024))
    return buf.value.decode(errors="replace")


def io_string_to_mat(s):
    T = np.zeros((4, 4))
    return T if lib().pgs_io_string_to_mat(s.encode(), T.ctypes.data_as(c_dp)) == 1 else None


def io_load_solved_posegraph(path):
    L = lib(); n = L.pgs_io_load_solved_posegraph(str(path).encode(), None, None, None, None, C.c_int32(0))
    if n < 0:
        raise PgsError(f"cannot load {path}")
    T = np.zeros((max(n, 1), 4, 4)); st = np.zeros(max(n, 1), np.int64); w = np.zeros(max(n, 1), np.int32); sid = np.zeros(max(n, 1), np.int32)
    L.pgs_io_load_solved_posegraph(str(path).encode(), T.ctypes.data_as(c_dp), st.ctypes.data_as(C.c_void_p), w.ctypes.data_as(c_ip), sid.ctypes.data_as(c_ip), C.c_int32(n))
    return T[:n], st[:n], w[:n], sid[:n]
