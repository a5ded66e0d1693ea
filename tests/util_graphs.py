"""Small seeded pose graphs built in numpy for the parity tests (independent of the C++ generator).
Returns plain arrays so the same problem can be loaded into the oracle and into libpgs."""
import numpy as np
from scipy.spatial.transform import Rotation as Rot


def _compose(qa, ta, qb, tb):
    Ra = Rot.from_quat(qa)
    return (Ra * Rot.from_quat(qb)).as_quat(), ta + Ra.apply(tb)


def _inv(q, t):
    Ri = Rot.from_quat(q).inv()
    return Ri.as_quat(), -Ri.apply(t)


def _canon(q):
    q = np.array(q, dtype=np.float64)
    if q.ndim == 1:
        return q if q[3] >= 0 else -q
    s = np.where(q[:, 3:4] >= 0, 1.0, -1.0)
    return q * s


def random_graph(n=60, fanout=3, n_loop=12, outlier_frac=0.0, seed=0, noise=1.0, reg=True, turn=0.01):
    """Gently turning 3-D walk, odometry edges (u,u-f) f=1..fanout with the reference's weight rule
    evaluated loosely (0.9^f), loop edges (a,b) with a>b, optional gross outliers."""
    rng = np.random.default_rng(seed)
    gq = np.zeros((n, 4)); gt = np.zeros((n, 3)); gq[0] = [0, 0, 0, 1]
    for i in range(1, n):
        dq = Rot.from_rotvec([0.002 * np.sin(i / 7.0), 0.003 * np.cos(i / 11.0), turn]).as_quat()
        gq[i], gt[i] = _compose(gq[i - 1], gt[i - 1], dq, np.array([1.0, 0.0, 0.0]))
    # odometry-integrated initial guess
    q0 = np.zeros((n, 4)); t0 = np.zeros((n, 3)); q0[0] = gq[0]; t0[0] = gt[0]
    for i in range(1, n):
        rq, rt = _compose(*_inv(gq[i - 1], gt[i - 1]), gq[i], gt[i])
        nq = Rot.from_rotvec(rng.normal(size=3) * 0.002 * noise).as_quat()
        rq, rt = _compose(rq, rt + rng.normal(size=3) * 0.02 * noise, nq, np.zeros(3))
        q0[i], t0[i] = _compose(q0[i - 1], t0[i - 1], rq, rt)
    oc1, oc2, oq, ot, ow = [], [], [], [], []
    for u in range(n):
        for f in range(1, fanout + 1):
            if u - f < 0:
                continue
            rq, rt = _compose(*_inv(q0[u], t0[u]), q0[u - f], t0[u - f])   # u_T_{u-f} from the "manager" poses
            oc1.append(u); oc2.append(u - f); oq.append(rq); ot.append(rt); ow.append(0.9 ** f)
    la, lb, lq, lt, lw, lout = [], [], [], [], [], []
    for _ in range(n_loop):
        b = int(rng.integers(0, n - 10)); a = int(rng.integers(b + 5, n))
        rq, rt = _compose(*_inv(gq[b], gt[b]), gq[a], gt[a])               # b_T_a
        out = rng.random() < outlier_frac
        if out:
            eq = Rot.from_rotvec(rng.normal(size=3)).as_quat(); et = rng.normal(size=3) * 8
        else:
            eq = Rot.from_rotvec(rng.normal(size=3) * 0.001 * noise).as_quat(); et = rng.normal(size=3) * 0.01 * noise
        rq, rt = _compose(rq, rt, eq, et)
        la.append(a); lb.append(b); lq.append(rq); lt.append(rt); lw.append(1.0); lout.append(out)
    g = dict(N=n, q=_canon(q0), t=t0, gt_q=_canon(gq), gt_t=gt,
             oc1=np.array(oc1, np.int32), oc2=np.array(oc2, np.int32), oq=_canon(np.array(oq).reshape(-1, 4)), ot=np.array(ot).reshape(-1, 3), ow=np.array(ow),
             la=np.array(la, np.int32), lb=np.array(lb, np.int32), lq=_canon(np.array(lq).reshape(-1, 4)), lt=np.array(lt).reshape(-1, 3), lw=np.array(lw),
             lout=np.array(lout, bool))
    if reg:
        g.update(rn=np.array([0], np.int32), rq=g["q"][:1].copy(), rt=g["t"][:1].copy(), rw=np.array([max(1.1, np.log(1 + n - 1) / 2.0)]))
    else:
        g.update(rn=np.zeros(0, np.int32), rq=np.zeros((0, 4)), rt=np.zeros((0, 3)), rw=np.zeros(0))
    return g


def load_oracle(g, switches=None):
    from oracle import pgo
    P = pgo.Problem()
    P.set_nodes(g["q"], g["t"])
    if len(g["oc1"]):
        P.add_odom_edges(g["oc1"], g["oc2"], g["oq"], g["ot"], g["ow"])
    if len(g["la"]):
        # reference binds a loop edge (a,b) as (c1,c2) = (b,a)  [PoseGraphSLAM.cpp:1553-1554]
        P.add_loop_edges(g["lb"], g["la"], g["lq"], g["lt"], g["lw"], s_init=switches)
    if len(g["rn"]):
        P.set_regularizers(g["rn"], g["rq"], g["rt"], g["rw"])
    return P


def load_pgs(g, switches=None, **opt):
    import solve_keyframe_pose_graph_b200 as pgs
    S = pgs.PoseGraphSolver(**opt)
    S.set_nodes(g["q"], g["t"])
    if len(g["oc1"]):
        S.add_odom_edges(g["oc1"], g["oc2"], g["oq"], g["ot"], g["ow"])
    if len(g["la"]):
        S.add_loop_edges(g["la"], g["lb"], g["lq"], g["lt"], g["lw"])
        if switches is not None:
            S.set_switches(switches)
    if len(g["rn"]):
        S.set_regularizers(g["rn"], g["rq"], g["rt"], g["rw"])
    return S


def rot_angle_between(qa, qb):
    """Geodesic angle (rad) between unit quaternion arrays (sign-insensitive).  Through the vector part of conj(qa) * qb
    (sin of the half angle), which keeps its precision for tiny angles where arccos of the dot product resolves ~3e-8."""
    qa = np.atleast_2d(qa); qb = np.atleast_2d(qb)
    va, wa = -qa[:, :3], qa[:, 3:4]
    vb, wb = qb[:, :3], qb[:, 3:4]
    v = wa * vb + wb * va + np.cross(va, vb)
    w = (wa * wb)[:, 0] - np.sum(va * vb, axis=1)
    return 2 * np.arctan2(np.linalg.norm(v, axis=1), np.abs(w))
