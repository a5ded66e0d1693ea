"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned to the reference's own front-end code by tests/test_reference_frontend.py
(oracle/_ref/libref_frontend.so: the reference's NodeDataManager.cpp / Worlds.cpp / PoseGraphSLAM.cpp compiled unmodified
over oracle/shim/ and single-stepped) and tests/test_reference_sets.py; the solve it calls (pgo) is Ceres restated.

Python restatement of the reference's host-side graph-construction rules, i.e. everything
`PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_optimize6DOF` does before and after
`ceres::Solve` (reference src/PoseGraphSLAM.cpp:1287-1950), and of the world bookkeeping it
consults (src/NodeDataManager.cpp:1127-1304, src/Worlds.cpp:6-275, src/utils/DisjointSet.h:152-257,
src/utils/MyDirectionalGraph.h:51-89).  Pure-Python loops: use on small graphs only.
4x4 products use numpy; rotation<->quaternion goes through the C oracle's Eigen restatement.
"""
import math
from collections import deque

import numpy as np

from . import pgo


class DisjointSetForest:
    """src/utils/DisjointSet.h: union by rank, path compression; link(X,Y): rank[X]>rank[Y] ? parent[Y]=X
    : parent[X]=Y (tie -> rank[Y]++)  (:241-257)."""

    def __init__(self):
        self.parent, self.rank = {}, {}

    def add_element(self, x):
        self.parent[x] = x; self.rank[x] = 0

    def exists(self, x):
        return x in self.parent

    def element_count(self):
        return len(self.parent)

    def find_set(self, x):
        p = self.parent[x]
        if p != x:
            p = self.parent[x] = self.find_set(p)
        return p

    def union_sets(self, x, y):
        sx, sy = self.find_set(x), self.find_set(y)
        if sx == sy:
            return
        if self.rank[sx] > self.rank[sy]:
            self.parent[sy] = sx
        else:
            self.parent[sx] = sy
            if self.rank[sx] == self.rank[sy]:
                self.rank[sy] += 1

    def set_count(self):
        return len({self.find_set(x) for x in self.parent})


def bfs_parents(n_vertices, edges, root):
    """MyDirectionalGraph::BFS (src/utils/MyDirectionalGraph.h:51-72): returns (parent, visited)."""
    adj = [[] for _ in range(n_vertices)]
    for v, w in edges:
        adj[v].append(w)
    parent = [-1] * n_vertices; visited = [False] * n_vertices
    q = deque([root]); visited[root] = True; parent[root] = -2
    while q:
        s = q.popleft()
        for i in adj[s]:
            if not visited[i]:
                visited[i] = True; parent[i] = s; q.append(i)
    return parent, visited


def path_from(parent, visited, v):
    """MyDirectionalGraph::get_path_from (:75-89), at most 100 hops."""
    if not visited[v]:
        return []
    out = []
    for _ in range(100):
        out.append(v)
        if parent[v] == -2:
            break
        v = parent[v]
    return out


class Worlds:
    """src/Worlds.cpp."""

    def __init__(self):
        self.rel = {}          # (m,n) -> m_T_n, iterated in sorted key order like std::map
        self.ds = DisjointSetForest()
        self.starts = []

    def n_worlds(self):
        return self.ds.element_count()

    def world_starts(self, t):
        self.starts.append(t); self.ds.add_element(len(self.starts) - 1)

    def find_setID_of_world_i(self, i):
        return self.ds.find_set(i) if self.ds.exists(i) else -1

    def is_exist(self, m, n):
        if m < 0 or n < 0:
            return False
        if m == n:
            return True
        if m >= self.n_worlds() or n >= self.n_worlds():
            return False
        return self.ds.find_set(m) == self.ds.find_set(n)

    def setPoseBetweenWorlds(self, m, n, T):
        self.rel[(m, n)] = np.array(T, dtype=np.float64)
        self.ds.union_sets(max(m, n), min(m, n))   # Worlds.cpp:168

    def getWorld2SetIDMap(self):
        return {w: self.find_setID_of_world_i(w) for w in range(self.n_worlds())}

    def getPoseBetweenWorlds(self, m, n):
        if m == n:
            return np.eye(4)
        if not self.is_exist(m, n):
            raise RuntimeError("reference exit(5): worlds not in the same set")
        if (m, n) in self.rel:
            return self.rel[(m, n)]
        if (n, m) in self.rel:
            return np.linalg.inv(self.rel[(n, m)])
        setid = self.ds.find_set(m)
        edges = []
        for (a, b) in sorted(self.rel.keys()):
            if self.ds.find_set(a) == setid and self.ds.find_set(b) == setid:
                edges.append((a, b)); edges.append((b, a))
        parent, visited = bfs_parents(self.n_worlds(), edges, n)
        path = path_from(parent, visited, m)
        ans = np.eye(4)
        for h in range(len(path) - 1):
            if (path[h], path[h + 1]) in self.rel:
                ans = ans @ self.rel[(path[h], path[h + 1])]
            elif (path[h + 1], path[h]) in self.rel:
                ans = ans @ np.linalg.inv(self.rel[(path[h + 1], path[h])])
            else:
                raise RuntimeError("reference exit(2)")
        self.setPoseBetweenWorlds(path[0], path[-1], ans)   # memoised (Worlds.cpp:137); the reference then falls off the end
        return ans


class Manager:
    """The slice of NodeDataManager the solver reads."""

    def __init__(self):
        self.stamps, self.poses = [], []
        self.edges, self.edge_pose, self.edge_w = [], [], []
        self.kidnap_starts, self.kidnap_ends = [], []
        self.kidnapped = False
        self.worlds = Worlds()

    def add_node(self, stamp, q, t):
        self.stamps.append(int(stamp)); self.poses.append(pgo.pose_to_mat4(q, t))
        if len(self.poses) == 1:
            self.worlds.world_starts(int(stamp))

    def add_loop_edge(self, a, b, q, t, w):
        self.edges.append((int(a), int(b))); self.edge_pose.append(pgo.pose_to_mat4(q, t)); self.edge_w.append(float(w))

    def kidnap_indicator(self, stamp, kidnapped):
        if kidnapped:
            self.kidnapped = True; self.kidnap_starts.append(int(stamp))
        else:
            self.kidnapped = False; self.kidnap_ends.append(int(stamp)); self.worlds.world_starts(int(stamp))

    def ingest(self, g):
        ev = sorted([(int(s), 1) for s in g["k0"]] + [(int(s), 0) for s in g["k1"]])
        pos = 0
        for stamp, kid in ev:
            while pos < g["N"] and g["stamps"][pos] <= stamp:
                self.add_node(g["stamps"][pos], g["q"][pos], g["t"][pos]); pos += 1
            self.kidnap_indicator(stamp, kid)
        while pos < g["N"]:
            self.add_node(g["stamps"][pos], g["q"][pos], g["t"][pos]); pos += 1
        for e in range(len(g["la"])):
            self.add_loop_edge(g["la"][e], g["lb"][e], g["lq"][e], g["lt"][e], g["lw"][e])

    def n_worlds(self):
        return len(self.kidnap_ends) + 1

    def which_world_is_this(self, t):
        """src/NodeDataManager.cpp:1127-1198, literally."""
        ks, ke = self.kidnap_starts, self.kidnap_ends
        if len(ks) == 0:
            return 0
        if len(ks) == 1:
            if t < ks[0]:
                return 0
            if len(ke) == 0:
                return -1
            return -1 if (ks[0] <= t <= ke[0]) else 1
        prev = 0
        if len(ks) == len(ke):
            for i in range(len(ks)):
                if prev < t <= ks[i]:
                    return i
                if ks[i] < t <= ke[i]:
                    return -(i + 1)
                prev = ke[i]
            return len(ke)
        for i in range(len(ks) - 1):
            if prev < t <= ks[i]:
                return i
            if ks[i] < t <= ke[i]:
                return -(i + 1)
            prev = ke[i]
        i = len(ks) - 1
        if ke[i - 1] < t <= ks[i]:
            return i
        return -(i + 1)

    def find_indexof_node(self, stamp):
        for i, s in enumerate(self.stamps):   # first node within 1 ms (NodeDataManager.cpp:274-299)
            if abs(s - stamp) < 1000000:
                return i
        return -1

    def nodeidx_of_world_i_started(self, i):
        if i < 0:
            return -3
        if i == 0:
            return 0
        if i - 1 < len(self.kidnap_ends):
            for r, s in enumerate(self.stamps):
                if self.which_world_is_this(s) == i:
                    return r
        return -4

    def nodeidx_of_world_i_ended(self, i):
        if i < 0 or i > len(self.kidnap_ends):
            return -1
        if i < len(self.kidnap_starts):
            return self.find_indexof_node(self.kidnap_starts[i])
        return len(self.stamps) - 1


class ReferenceFrontEnd:
    """State of the reference's solver thread across triggers + one `trigger()` = one wake-up."""

    def __init__(self, manager, odom_fanout=5, options=None):
        self.m = manager
        self.fanout = odom_fanout
        self.options = options
        self.opt_q, self.opt_t, self.opt_s = [], [], []
        self.solved_until = 0
        self.n_constant = 0
        self.prev_loopedge_len = 0
        self.changes = {}
        # accumulated residual blocks
        self.odom = []     # (u, u-f, q, t, w)
        self.loops = []    # (e, a, b, q, t, w)
        self.regs = []     # (node, q, t, w)

    def pose(self, i):
        return pgo.pose_to_mat4(self.opt_q[i], self.opt_t[i])

    def _set(self, i, T):
        q, t = pgo.mat4_to_pose(T)
        self.opt_q[i] = q; self.opt_t[i] = t

    def load_state(self):
        """PoseGraphSLAM::load_state (src/PoseGraphSLAM.cpp:40-170): every keyframe already in the manager becomes a
        CONSTANT optimisation variable at ws_T_w * w_T_c; solvedUntil moves to the last one."""
        m = self.m; W = m.worlds
        for yp in range(len(self.opt_q), len(m.poses)):
            world = m.which_world_is_this(m.stamps[yp]); setid = W.find_setID_of_world_i(world)
            ws_T_w = np.eye(4)
            if world >= 0 and world != setid:                      # :104-116
                if not W.is_exist(setid, world):
                    raise RuntimeError("reference exit(1)")
                ws_T_w = W.getPoseBetweenWorlds(setid, world)
            q, t = pgo.mat4_to_pose(ws_T_w @ m.poses[yp])           # :118-131
            self.opt_q.append(q); self.opt_t.append(t)
        self.n_constant = len(m.poses)                              # :150-151
        self.solved_until = len(m.poses) - 1                        # :165

    def trigger(self, solve=True):
        m = self.m
        node_len, loopedge_len = len(m.poses), len(m.edges)
        if self.prev_loopedge_len == loopedge_len or m.kidnapped:
            return None
        W = m.worlds
        while len(self.opt_q) < node_len:                       # :1340-1355
            self.opt_q.append(np.array([0, 0, 0, 1.0])); self.opt_t.append(np.zeros(3))
        while len(self.opt_s) < loopedge_len:                   # :1359-1367
            self.opt_s.append(0.99)
        ww = lambda i: m.which_world_is_this(m.stamps[i])
        # ---- loop edges (:1381-1559)
        for e in range(self.prev_loopedge_len, loopedge_len):
            a, b = m.edges[e]; bTa = m.edge_pose[e]
            wa, wb = ww(a), ww(b)
            if wa < 0 or wb < 0:
                continue
            if wa != wb and not W.is_exist(wb, wa):
                wb_T_wa = (m.poses[b] @ bTa) @ pgo.inv4(m.poses[a])
                before = W.getWorld2SetIDMap()
                W.setPoseBetweenWorlds(wb, wa, wb_T_wa)
                after = W.getWorld2SetIDMap()
                self.changes = {k: (v, after[k]) for k, v in before.items() if v != after[k]}
            q, t = pgo.mat4_to_pose(bTa)
            self.loops.append((e, a, b, q, t, m.edge_w[e]))
        # ---- odometry edges (:1570-1639)
        for u in range(self.solved_until + 1, node_len):
            su = W.find_setID_of_world_i(ww(u))
            for f in range(1, self.fanout + 1):
                wumf = ww(u - f) if u - f >= 0 else -1
                sumf = W.find_setID_of_world_i(wumf)
                if su < 0 or sumf < 0 or u - f < 0:
                    continue
                u_M_umf = pgo.inv4(m.poses[u]) @ m.poses[u - f]
                yaw = pgo.r2ypr_deg(u_M_umf)[0]
                w = math.pow(0.9, f) * math.exp(-yaw * yaw / 6.0)
                q, t = pgo.mat4_to_pose(u_M_umf)
                self.odom.append((u, u - f, q, t, w))
        # ---- initial guesses (:1649-1793)
        s_u = self.solved_until
        s_w = ww(s_u)
        if s_w < 0:
            s_w = -s_w - 1
        for u in range(node_len):
            wu = ww(u); setid = W.find_setID_of_world_i(wu)
            if setid < 0:
                continue
            wset_T_w = np.eye(4)
            if setid != wu:
                if not W.is_exist(setid, wu):
                    raise RuntimeError("reference exit(3)")
                wset_T_w = W.getPoseBetweenWorlds(setid, wu)
            before = u <= s_u
            in_change = wu in self.changes
            if in_change and before:
                if setid == s_w:
                    raise RuntimeError("reference exit(8)")
                old, new = self.changes[wu]
                self._set(u, W.getPoseBetweenWorlds(new, old) @ self.pose(u))
            elif not before:
                if s_w == wu:
                    self._set(u, self.pose(s_u) @ (pgo.inv4(m.poses[s_u]) @ m.poses[u]))
                else:
                    self._set(u, wset_T_w @ m.poses[u])
            elif s_u == 0:
                self._set(u, m.poses[u])
        # ---- regularisers (:1801-1879)
        self.regs = []
        for w in range(m.n_worlds()):
            setid = W.find_setID_of_world_i(w)
            start, end = m.nodeidx_of_world_i_started(w), m.nodeidx_of_world_i_ended(w)
            if start < 0:
                continue
            if setid >= 0 and setid == w:
                arg = 1 + end - start
                x = math.log(arg) / 2.0 if arg > 0 else float("nan")
                weight = x if 1.1 < x else 1.1          # std::max(1.1, x)
                self.regs.append((start, np.array(self.opt_q[start]), np.array(self.opt_t[start]), weight))
        self.changes = {}
        summary = None
        if solve:
            summary = self.solve()
        self.solved_until = node_len - 1                        # :1908
        self.prev_loopedge_len = loopedge_len
        return summary

    def alternative_terms(self, kind):
        """The residual blocks the reference's SWITCHED-OFF builds add for the same session, as keyword arguments of
        pgo.fourdof_eval: kind 0 FourDOFError::Create(u_M_umf, odom_edge_weight) on (u, u-f) (the commented call at
        src/PoseGraphSLAM.cpp:1630); kind 1 FourDOFErrorWithSwitchingConstraints::Create(bTa, weight) on
        (paur.second, paur.first, switch e) (:1551); kind 2 the __USE_YPR_REP build: (ypr, t) variables (:228-247),
        QinFourDOFWeightError(u_M_umf translation, yaw of u_M_umf, pitch and roll of w_M_u) on (u, u-f) (:1608-1626), then
        QinFourDOFWeightError(bTa translation, yaw of bTa, pitch and roll of the manager pose of paur.first) on
        (paur.second, paur.first) (:1389-1392,1534-1548)."""
        m = self.m
        t = np.array(self.opt_t).reshape(-1, 3)
        if kind == 0:
            return dict(rot=np.array(self.opt_q).reshape(-1, 4), t=t, c1=[o[0] for o in self.odom], c2=[o[1] for o in self.odom],
                        obs_rot=np.array([o[2] for o in self.odom]).reshape(-1, 4), obs_t=np.array([o[3] for o in self.odom]).reshape(-1, 3),
                        weight=[o[4] for o in self.odom], sw=None)
        if kind == 1:
            return dict(rot=np.array(self.opt_q).reshape(-1, 4), t=t, c1=[l[2] for l in self.loops], c2=[l[1] for l in self.loops],
                        obs_rot=np.array([l[3] for l in self.loops]).reshape(-1, 4), obs_t=np.array([l[4] for l in self.loops]).reshape(-1, 3),
                        weight=[l[5] for l in self.loops], sw=[self.opt_s[l[0]] for l in self.loops])
        ypr = np.array([pgo.r2ypr_deg(self.pose(i)) for i in range(len(self.opt_q))]).reshape(-1, 3)   # eigenmat_to_rawyprt
        c1, c2, obs_rot, obs_t = [], [], [], []
        for (u, umf, q, tt, w) in self.odom:
            u_M_umf = pgo.pose_to_mat4(q, tt); own = pgo.r2ypr_deg(m.poses[u])
            c1.append(u); c2.append(umf); obs_t.append(u_M_umf[:3, 3]); obs_rot.append([pgo.r2ypr_deg(u_M_umf)[0], own[1], own[2]])
        for (e, a, b, q, tt, w) in self.loops:
            bTa = m.edge_pose[e]; own = pgo.r2ypr_deg(m.poses[a])
            c1.append(b); c2.append(a); obs_t.append(bTa[:3, 3]); obs_rot.append([pgo.r2ypr_deg(bTa)[0], own[1], own[2]])
        return dict(rot=ypr, t=t, c1=c1, c2=c2, obs_rot=np.array(obs_rot).reshape(-1, 3), obs_t=np.array(obs_t).reshape(-1, 3), weight=None, sw=None)

    def problem(self):
        P = pgo.Problem()
        P.set_nodes(np.array(self.opt_q), np.array(self.opt_t))
        if self.n_constant:
            P.set_constant_nodes(0, self.n_constant)
        if self.odom:
            P.add_odom_edges([o[0] for o in self.odom], [o[1] for o in self.odom], np.array([o[2] for o in self.odom]),
                             np.array([o[3] for o in self.odom]), [o[4] for o in self.odom])
        if self.loops:
            # bound as (b, a, switch e)  (:1553-1555)
            P.add_loop_edges([l[2] for l in self.loops], [l[1] for l in self.loops], np.array([l[3] for l in self.loops]),
                             np.array([l[4] for l in self.loops]), [l[5] for l in self.loops], s_init=[self.opt_s[l[0]] for l in self.loops])
        if self.regs:
            P.set_regularizers([r[0] for r in self.regs], np.array([r[1] for r in self.regs]), np.array([r[2] for r in self.regs]), [r[3] for r in self.regs])
        return P

    def solve(self):
        P = self.problem()
        s = P.solve(self.options)
        q, t = P.poses(); sw = P.switches()
        self.opt_q = [q[i] for i in range(len(q))]; self.opt_t = [t[i] for i in range(len(t))]
        for k, l in enumerate(self.loops):
            self.opt_s[l[0]] = sw[k]
        return s
