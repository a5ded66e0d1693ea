"""ctypes binding of include/pgs.h.  Mirrors the C-ABI one to one; `PoseGraphSolver` is the
RAII convenience used by tests and bench.py.  Loading fails loudly if libpgs.so is missing —
the product never falls back to a CPU implementation."""
import ctypes as C
import os
import re

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class PgsError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("jacobi_scaling", C.c_int32),
        ("switch_init", C.c_double),
        ("device", C.c_int32),
        ("linear_solver", C.c_int32),
        ("pcg_max_iterations", C.c_int32),
        ("pcg_tolerance", C.c_double),
        ("chains", C.c_int32),
        ("check_linear_solves", C.c_int32),
        ("max_factor_bytes", C.c_double),
        ("max_factor_flops", C.c_double),
    ]


class Iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32),
        ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double),
        ("step_norm", C.c_double), ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double),
        ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32), ("linear_solver_iterations", C.c_int32),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("termination", C.c_int32), ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
        ("num_iterations", C.c_int32), ("linear_solver_iterations", C.c_int32),
        ("ms_sweep", C.c_double), ("ms_assemble", C.c_double), ("ms_linear_solve", C.c_double), ("ms_total", C.c_double),
        ("factor_nnz", C.c_int64),
        ("linear_solver_used", C.c_int32), ("n_chains", C.c_int32), ("factor_flops", C.c_double),
        ("max_linear_backward_error", C.c_double), ("fixed_cost", C.c_double), ("ms_comm", C.c_double),
    ]


class Sizes(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_odom", C.c_int32), ("n_loop", C.c_int32), ("n_reg", C.c_int32), ("n_pairs", C.c_int32)]


class DistStats(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("n_interior_nodes", C.c_int32), ("n_border_nodes", C.c_int32),
                ("n_odom_owned", C.c_int32), ("n_loop_owned", C.c_int32), ("n_reg_owned", C.c_int32),
                ("border_buffer_bytes", C.c_int64), ("n_collectives", C.c_int64), ("bytes_reduced", C.c_int64),
                ("n_chains", C.c_int32), ("n_local_border_nodes", C.c_int32), ("factor_nnz", C.c_int64), ("ms_comm", C.c_double),
                ("ms_eliminate", C.c_double), ("ms_exchange", C.c_double), ("ms_border", C.c_double)]


SKYLINE_CHOLESKY, BLOCK_PCG = 0, 1
TERMINATION = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}


def library_path():
    return os.path.join(_PKG, "libpgs.so")


def exported_symbols():
    """Function names declared in include/pgs.h (what the shared library must export)."""
    txt = open(os.path.join(_ROOT, "include", "pgs.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pgs_[a-z_0-9]+)\s*\(", txt)))


def lib():
    global _LIB
    if _LIB is None:
        p = library_path()
        if not os.path.exists(p):
            raise PgsError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C solve_keyframe_pose_graph_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(p)
        L.pgs_last_error.restype = C.c_char_p
        L.pgs_last_error.argtypes = [C.c_void_p]
        L.pgs_sweep_algorithmic_bytes.restype = C.c_int64
        L.pgs_sweep_algorithmic_bytes.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _d(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def dist_unique_id():
    """128-byte NCCL unique id (call on rank 0, broadcast to the other ranks)."""
    buf = C.create_string_buffer(128)
    rc = lib().pgs_dist_unique_id(buf)
    if rc != 0:
        raise PgsError(f"pgs_dist_unique_id failed ({rc}): {lib().pgs_last_error(None).decode(errors="replace")}")
    return bytes(buf.raw)


def partition(n_nodes, world, oc1, oc2, la, lb, rn, chains_per_rank=0):
    """Host-only view of the node-range plan (include/pgs.h pgs_partition)."""
    oc1, p1 = _i(oc1); oc2, p2 = _i(oc2); la, pa = _i(la); lb, pb = _i(lb); rn, pr = _i(rn)
    node_owner = np.zeros(max(n_nodes, 1), np.int32); oo = np.zeros(max(len(oc1), 1), np.int32)
    lo = np.zeros(max(len(la), 1), np.int32); ro = np.zeros(max(len(rn), 1), np.int32); nb = C.c_int32(0)
    cut = np.zeros(world + 1, np.int32); node_chain = np.zeros(max(n_nodes, 1), np.int32)
    down = np.zeros(world * max(chains_per_rank, 2), np.int32); nc = C.c_int32(0)
    rc = lib().pgs_partition(C.c_int32(n_nodes), C.c_int32(world), C.c_int32(len(oc1)), p1, p2, C.c_int32(len(la)), pa, pb, C.c_int32(len(rn)), pr,
                             node_owner.ctypes.data_as(c_ip), oo.ctypes.data_as(c_ip), lo.ctypes.data_as(c_ip), ro.ctypes.data_as(c_ip), C.byref(nb),
                             C.c_int32(chains_per_rank), cut.ctypes.data_as(c_ip), node_chain.ctypes.data_as(c_ip), down.ctypes.data_as(c_ip), C.byref(nc))
    if rc != 0:
        raise PgsError(f"pgs_partition failed ({rc})")
    return dict(node_owner=node_owner[:n_nodes], odom_owner=oo[:len(oc1)], loop_owner=lo[:len(la)], reg_owner=ro[:len(rn)], n_border=nb.value,
                cut=cut, node_chain=node_chain[:n_nodes], chain_down=down[:nc.value], n_chains=nc.value)


class ComposeInput(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("mgr_T", c_dp), ("world_id", c_ip), ("n_slam", C.c_int32), ("slam_q", c_dp), ("slam_t", c_dp),
                ("solved_until", C.c_int32), ("solved_until_world", C.c_int32), ("n_worlds", C.c_int32), ("world_end", c_ip),
                ("world_setid", c_ip), ("ws_exists", C.POINTER(C.c_uint8)), ("ws_T_w", c_dp)]


def compose_poses(mgr_T, world_id, slam_q, slam_t, solved_until, solved_until_world, world_end, world_setid, ws_exists, ws_T_w, device=0):
    """include/pgs_compose.h in one call: the Composer pose assembly (reference src/Composer.cpp:24-209) on the device.
    Returns (T [n,4,4], kernel ms, total ms)."""
    L = lib()
    L.pgs_compose_last_error.restype = C.c_char_p; L.pgs_compose_last_error.argtypes = [C.c_void_p]
    h = C.c_void_p()
    if L.pgs_compose_create(C.c_int32(device), C.byref(h)) != 0:
        raise PgsError("pgs_compose_create failed: no usable CUDA device (there is no CPU fallback)")
    try:
        mgr_T, pm = _d(np.asarray(mgr_T).reshape(-1, 16)); wid, pw = _i(world_id)
        sq, pq = _d(np.asarray(slam_q).reshape(-1, 4)); st, pt = _d(np.asarray(slam_t).reshape(-1, 3))
        we, pe = _i(world_end); ws, ps = _i(world_setid)
        ex = np.ascontiguousarray(ws_exists, dtype=np.uint8); wt, pwt = _d(np.asarray(ws_T_w).reshape(-1, 16))
        inp = ComposeInput(len(wid), pm, pw, len(sq), pq, pt, int(solved_until), int(solved_until_world), len(we), pe, ps,
                           ex.ctypes.data_as(C.POINTER(C.c_uint8)), pwt)
        out = np.zeros((max(len(wid), 1), 4, 4))
        rc = L.pgs_compose_run(h, C.byref(inp), out.ctypes.data_as(c_dp))
        if rc != 0:
            raise PgsError(f"pgs_compose_run failed ({rc}): {L.pgs_compose_last_error(h).decode(errors="replace")}")
        a = C.c_double(0); b = C.c_double(0)
        L.pgs_compose_last_timing(h, C.byref(a), C.byref(b))
        return out[: len(wid)], a.value, b.value
    finally:
        L.pgs_compose_destroy(h)


class FourDofInput(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_nodes", C.c_int32), ("rot", c_dp), ("t", c_dp), ("n_edges", C.c_int32), ("c1", c_ip), ("c2", c_ip),
                ("obs_rot", c_dp), ("obs_t", c_dp), ("weight", c_dp), ("sw", c_dp)]


FOURDOF_ERROR, FOURDOF_SWITCH, FOURDOF_QIN = 0, 1, 2
FOURDOF_SHAPES = {0: (6, 12, 4), 1: (7, 13, 4), 2: (4, 8, 3)}   # kind -> residual rows, tangent columns, doubles per rotation record


def fourdof_evaluate(kind, rot, t, c1, c2, obs_rot, obs_t, weight=None, sw=None, jac=True, device=0):
    """include/pgs_fourdof.h in one call: the reference's alternative (switched-off) edge functors
    (src/CeresResidues.h:252-546) over an edge list, on the device.  Returns dict(cost, r[E,NR], J[E,NR,NC], ms_kernel)."""
    L = lib()
    L.pgs_fourdof_last_error.restype = C.c_char_p; L.pgs_fourdof_last_error.argtypes = [C.c_void_p]
    NR, NC, RW = FOURDOF_SHAPES[kind]
    h = C.c_void_p()
    if L.pgs_fourdof_create(C.c_int32(device), C.byref(h)) != 0:
        raise PgsError("pgs_fourdof_create failed: no usable CUDA device (there is no CPU fallback)")
    try:
        rot, rp = _d(np.asarray(rot).reshape(-1, RW)); t, tp = _d(np.asarray(t).reshape(-1, 3))
        c1, p1 = _i(c1); c2, p2 = _i(c2)
        E = len(c1)
        obs_rot, orp = _d(np.asarray(obs_rot).reshape(E, RW)); obs_t, otp = _d(np.asarray(obs_t).reshape(E, 3))
        weight, wp = _d(weight); sw, sp = _d(sw)
        inp = FourDofInput(kind, len(rot), rp, tp, E, p1, p2, orp, otp, wp, sp)
        r = np.zeros((E, NR)); J = np.zeros((E, NR, NC)) if jac else None
        cost = C.c_double(0)
        rc = L.pgs_fourdof_evaluate(h, C.byref(inp), r.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp) if jac else None, C.byref(cost))
        if rc != 0:
            raise PgsError(f"pgs_fourdof_evaluate failed ({rc}): {L.pgs_fourdof_last_error(h).decode(errors="replace")}")
        ms = C.c_double(0)
        L.pgs_fourdof_last_timing(h, C.byref(ms))
        return dict(cost=cost.value, r=r, J=J, ms_kernel=ms.value)
    finally:
        L.pgs_fourdof_destroy(h)


def default_options(**kw):
    o = Options()
    lib().pgs_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class PoseGraphSolver:
    def __init__(self, options=None, **kw):
        self.L = lib()
        self.opt = options or default_options(**kw)
        self.h = C.c_void_p()
        rc = self.L.pgs_create(C.byref(self.opt), C.byref(self.h))
        if rc != 0:
            raise PgsError(f"pgs_create failed ({rc}): {self.L.pgs_last_error(None).decode(errors="replace")}")
        self.N = 0
        self.n_odom = self.n_loop = self.n_reg = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.pgs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PgsError(f"pgs error {rc}: {self.L.pgs_last_error(self.h).decode(errors="replace")}")

    # ---- construction
    def set_nodes(self, q, t):
        q, qp = _d(q); t, tp = _d(t)
        self._ck(self.L.pgs_set_nodes(self.h, C.c_int32(q.shape[0]), qp, tp)); self.N = q.shape[0]

    def append_nodes(self, q, t):
        q, qp = _d(q); t, tp = _d(t)
        self._ck(self.L.pgs_append_nodes(self.h, C.c_int32(q.shape[0]), qp, tp)); self.N += q.shape[0]

    def update_nodes(self, first, q, t):
        q, qp = _d(q); t, tp = _d(t)
        self._ck(self.L.pgs_update_nodes(self.h, C.c_int32(first), C.c_int32(q.shape[0]), qp, tp))

    def add_odom_edges(self, c1, c2, q, t, w):
        c1, c1p = _i(c1); c2, c2p = _i(c2); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        self._ck(self.L.pgs_add_odom_edges(self.h, C.c_int32(len(c1)), c1p, c2p, qp, tp, wp)); self.n_odom += len(c1)

    def add_loop_edges(self, a, b, q, t, w):
        a, ap = _i(a); b, bp = _i(b); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        self._ck(self.L.pgs_add_loop_edges(self.h, C.c_int32(len(a)), ap, bp, qp, tp, wp)); self.n_loop += len(a)

    def set_regularizers(self, node, q, t, w):
        node, np_ = _i(node); q, qp = _d(q); t, tp = _d(t); w, wp = _d(w)
        self._ck(self.L.pgs_set_regularizers(self.h, C.c_int32(len(node)), np_, qp, tp, wp)); self.n_reg = len(node)

    def set_constant_nodes(self, first, n, constant=True):
        self._ck(self.L.pgs_set_constant_nodes(self.h, C.c_int32(first), C.c_int32(n), C.c_int32(int(constant))))

    def set_switches(self, s, first=0):
        s, sp = _d(s)
        self._ck(self.L.pgs_set_switches(self.h, C.c_int32(first), C.c_int32(len(s)), sp))

    # ---- getters
    def poses(self, first=0, n=None):
        n = self.N - first if n is None else n
        q = np.empty((n, 4)); t = np.empty((n, 3))
        self._ck(self.L.pgs_get_poses(self.h, C.c_int32(first), C.c_int32(n), q.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp)))
        return q, t

    def switches(self):
        s = np.empty(self.n_loop)
        if self.n_loop:
            self._ck(self.L.pgs_get_switches(self.h, C.c_int32(0), C.c_int32(self.n_loop), s.ctypes.data_as(c_dp)))
        return s

    def sizes(self):
        s = Sizes(); self._ck(self.L.pgs_get_sizes(self.h, C.byref(s))); return s

    # ---- evaluation
    def evaluate(self, jac=True, residuals=True):
        cost = C.c_double(0)
        out = {}
        ptr = lambda a: a.ctypes.data_as(c_dp)
        if residuals:
            out.update(r_o=np.zeros((self.n_odom, 6)), r_l=np.zeros((self.n_loop, 7)), r_r=np.zeros((self.n_reg, 6)))
        if jac:
            out.update(J_o=np.zeros((self.n_odom, 6, 12)), J_l=np.zeros((self.n_loop, 7, 13)), J_r=np.zeros((self.n_reg, 6, 6)))
        g = lambda k: ptr(out[k]) if k in out else None
        self._ck(self.L.pgs_evaluate(self.h, C.byref(cost), g("r_o"), g("J_o"), g("r_l"), g("J_l"), g("r_r"), g("J_r")))
        out["cost"] = cost.value
        return out

    def gradient(self):
        gp = np.zeros((self.N, 6)); gs = np.zeros(max(self.n_loop, 1))
        self._ck(self.L.pgs_gradient(self.h, gp.ctypes.data_as(c_dp), gs.ctypes.data_as(c_dp)))
        return gp, gs[: self.n_loop]

    def assemble(self):
        sz = self.sizes(); P = sz.n_pairs
        diag = np.zeros((self.N, 6, 6)); hi = np.zeros(max(P, 1), np.int32); lo = np.zeros(max(P, 1), np.int32)
        off = np.zeros((max(P, 1), 6, 6)); lv = np.zeros((max(self.n_loop, 1), 12)); lh = np.zeros(max(self.n_loop, 1))
        self._ck(self.L.pgs_assemble(self.h, diag.ctypes.data_as(c_dp), hi.ctypes.data_as(c_ip), lo.ctypes.data_as(c_ip), off.ctypes.data_as(c_dp),
                                     lv.ctypes.data_as(c_dp), lh.ctypes.data_as(c_dp)))
        return dict(diag=diag, pair_hi=hi[:P], pair_lo=lo[:P], offdiag=off[:P], loop_v=lv[: self.n_loop], loop_hss=lh[: self.n_loop])

    def linear_step(self, radius):
        dp = np.zeros((self.N, 6)); ds = np.zeros(max(self.n_loop, 1)); mcc = C.c_double(0); it = C.c_int32(0)
        self._ck(self.L.pgs_linear_step(self.h, C.c_double(radius), dp.ctypes.data_as(c_dp), ds.ctypes.data_as(c_dp), C.byref(mcc), C.byref(it)))
        return dp, ds[: self.n_loop], mcc.value, it.value

    def solve(self):
        s = Summary(); cap = self.opt.max_num_iterations + 8
        its = (Iteration * cap)()
        self._ck(self.L.pgs_solve(self.h, C.byref(s), its, C.c_int32(cap)))
        rows = [{f: getattr(its[i], f) for f, _ in Iteration._fields_} for i in range(min(s.num_iterations, cap))]
        d = {f: getattr(s, f) for f, _ in Summary._fields_}
        d["termination"] = TERMINATION[s.termination]; d["iterations"] = rows
        return d

    # ---- multi-GPU
    def dist_init(self, rank, world, unique_id):
        self._ck(self.L.pgs_dist_init(self.h, C.c_int32(rank), C.c_int32(world), C.c_char_p(unique_id)))

    def dist_init_local(self, rank, world, group):
        """In-process transport: `world` handles of this process, one thread each, naming the same group."""
        self._ck(self.L.pgs_dist_init_local(self.h, C.c_int32(rank), C.c_int32(world), C.c_char_p(group.encode())))

    def linear_backward_errors(self):
        n = C.c_int32(0); out = np.zeros(256)
        self._ck(self.L.pgs_get_linear_backward_errors(self.h, out.ctypes.data_as(c_dp), C.c_int32(len(out)), C.byref(n)))
        return out[: min(n.value, len(out))].copy()

    def dist_stats(self):
        s = DistStats(); self._ck(self.L.pgs_dist_get_stats(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in DistStats._fields_}

    # ---- measurement hooks
    def time_sweep(self, mode=0, reps=10, flush_l2=False):
        """-> (ms per step = sweep kernel + cost reduction, ms of the sweep kernel alone, kernel launches)"""
        ms = C.c_double(0); msk = C.c_double(0); n = C.c_int64(0)
        self._ck(self.L.pgs_time_sweep(self.h, C.c_int32(mode), C.c_int32(reps), C.c_int32(int(flush_l2)), C.byref(ms), C.byref(msk), C.byref(n)))
        return ms.value, msk.value, n.value

    def time_stream_write(self, nbytes, reps=10, flush_l2=True):
        """ms of a pure streaming write of `nbytes`, timed like time_sweep (practical ceiling of a store-bound kernel)."""
        ms = C.c_double(0)
        self._ck(self.L.pgs_time_stream_write(self.h, C.c_int64(int(nbytes)), C.c_int32(reps), C.c_int32(int(flush_l2)), C.byref(ms)))
        return ms.value

    def evaluate_from_host_ptr(self, q_ptr, t_ptr, s_ptr):
        """Raw-pointer variant (pinned torch tensors): addresses as ints, 0 for 'reuse'."""
        cost = C.c_double(0)
        self._ck(self.L.pgs_evaluate_from_host(self.h, C.c_void_p(q_ptr), C.c_void_p(t_ptr), C.c_void_p(s_ptr), C.byref(cost)))
        return cost.value

    def evaluate_from_host(self, q, t, s=None):
        q, qp = _d(q); t, tp = _d(t); s, sp = _d(s)
        cost = C.c_double(0)
        self._ck(self.L.pgs_evaluate_from_host(self.h, qp, tp, sp, C.byref(cost)))
        return cost.value

    def sweep_bytes(self):
        return int(self.L.pgs_sweep_algorithmic_bytes(self.h))
