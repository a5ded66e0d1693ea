"""Set bookkeeping against the REFERENCE'S OWN CODE (SURVEY §8a row "World / set bookkeeping", §8c).

oracle/_ref/libref_sets.so is compiled from the reference's dependency-free headers where they lie under
/root/reference (src/utils/DisjointSet.h, src/utils/MyDirectionalGraph.h; recipe: oracle/Makefile, wrapper
oracle/ref_sets_capi.cpp).  The product's union-find and Worlds classes (csrc/host/DisjointSet.h, Worlds.cpp, compiled
for this test by tests/sets_hostcheck.cpp) and the oracle's Python front-end restatement are driven op by op against it:
which world becomes a set root — hence the frame every keyframe is initialised in and where regularisers go — and which
chain of relative poses an inferred world-to-world transform is built from are pinned to the real code, not to a
reading of it.  The library is built in the container that has the reference; without it the tests skip."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import frontend, pgo

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_sets.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_sets.so not built (needs /root/reference; make -C oracle)")


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(REF_SO)
    L.ref_dsf_create.restype = C.c_void_p; L.ref_graph_create.restype = C.c_void_p
    for f, nint in dict(ref_dsf_destroy=0, ref_dsf_element_count=0, ref_dsf_set_count=0, ref_graph_destroy=0, ref_dsf_add_element=1, ref_dsf_exists=1,
                        ref_dsf_find_set=1, ref_graph_bfs=1, ref_dsf_union_sets=2, ref_graph_add_edge=2).items():
        getattr(L, f).argtypes = [C.c_void_p] + [C.c_int] * nint
    L.ref_graph_get_path_from.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
    return L


@pytest.fixture(scope="module")
def ours(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sets") / "sets_hostcheck.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "sets_hostcheck.cpp")])
    L = C.CDLL(so)
    L.ours_dsf_create.restype = C.c_void_p; L.ours_worlds_create.restype = C.c_void_p
    for f in ("ours_dsf_destroy", "ours_dsf_element_count", "ours_dsf_set_count", "ours_worlds_destroy", "ours_worlds_n_keys"):
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("ours_dsf_add_element", "ours_dsf_exists", "ours_dsf_find_set", "ours_worlds_find_setid"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int]
    for f in ("ours_dsf_union_sets", "ours_worlds_is_exist"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ours_worlds_world_starts.argtypes = [C.c_void_p, C.c_longlong]
    L.ours_worlds_set_pose.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.ours_worlds_get_pose.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return L


def test_reference_disjoint_set_sample_of_its_own_test_file(ref):
    """src/test_disjointset.cpp:26-44 turned into assertions on the reference's class itself."""
    d = ref.ref_dsf_create()
    for i in range(5):
        ref.ref_dsf_add_element(d, i)
    assert ref.ref_dsf_element_count(d) == 5 and ref.ref_dsf_set_count(d) == 5
    ref.ref_dsf_union_sets(d, 3, 2)                       # tie: rank[3] == rank[2] -> parent[3] = 2, rank[2] = 1
    assert ref.ref_dsf_find_set(d, 3) == 2 and ref.ref_dsf_set_count(d) == 4
    ref.ref_dsf_union_sets(d, 2, 0)                       # rank[2] = 1 > rank[0] = 0 -> parent[0] = 2: the LARGER id stays root
    assert ref.ref_dsf_find_set(d, 0) == 2 and ref.ref_dsf_find_set(d, 3) == 2 and ref.ref_dsf_set_count(d) == 3
    assert ref.ref_dsf_find_set(d, 7) == -1 and not ref.ref_dsf_exists(d, 7)
    ref.ref_dsf_destroy(d)


@pytest.mark.parametrize("seed", range(6))
def test_union_find_matches_the_reference_class_op_by_op(ref, ours, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 40))
    R, O, P = ref.ref_dsf_create(), ours.ours_dsf_create(), frontend.DisjointSetForest()
    for i in range(n):
        ref.ref_dsf_add_element(R, i); ours.ours_dsf_add_element(O, i); P.add_element(i)
    for _ in range(4 * n):
        op = rng.integers(0, 3)
        x, y = int(rng.integers(0, n)), int(rng.integers(0, n))
        if op == 0:                                          # the call pattern of Worlds::setPoseBetweenWorlds (Worlds.cpp:167)
            a, b = max(x, y), min(x, y)
            ref.ref_dsf_union_sets(R, a, b); ours.ours_dsf_union_sets(O, a, b); P.union_sets(a, b)
        elif op == 1:                                        # arbitrary order too
            ref.ref_dsf_union_sets(R, x, y); ours.ours_dsf_union_sets(O, x, y); P.union_sets(x, y)
        else:                                                # finds compress paths: part of the state
            assert ref.ref_dsf_find_set(R, x) == ours.ours_dsf_find_set(O, x) == P.find_set(x)
        assert ref.ref_dsf_set_count(R) == ours.ours_dsf_set_count(O) == len({P.find_set(i) for i in range(n)})
    roots = [ref.ref_dsf_find_set(R, i) for i in range(n)]
    assert roots == [ours.ours_dsf_find_set(O, i) for i in range(n)] == [P.find_set(i) for i in range(n)]
    ref.ref_dsf_destroy(R); ours.ours_dsf_destroy(O)


@pytest.mark.parametrize("seed", range(4))
def test_bfs_parents_and_paths_match_the_reference_graph(ref, seed):
    rng = np.random.default_rng(100 + seed)
    V = int(rng.integers(4, 25))
    edges = [(int(rng.integers(0, V)), int(rng.integers(0, V))) for _ in range(int(rng.integers(V, 3 * V)))]
    G = ref.ref_graph_create(V)
    for v, w in edges:
        ref.ref_graph_add_edge(G, v, w)
    s = int(rng.integers(0, V))
    ref.ref_graph_bfs(G, s)
    parent, visited = frontend.bfs_parents(V, edges, s)
    buf = (C.c_int * 128)()
    for v in range(V):
        k = ref.ref_graph_get_path_from(G, v, buf, 128)
        assert list(buf[:k]) == frontend.path_from(parent, visited, v)
    ref.ref_graph_destroy(G)


def _rand_pose(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return pgo.pose_to_mat4(q, rng.normal(size=3) * 10)


class ReferenceWorldsModel:
    """Worlds::setPoseBetweenWorlds / is_exist / getPoseBetweenWorlds (src/Worlds.cpp:6-195) with the REFERENCE'S classes
    doing the union-find and the breadth-first search; only the map and the matrix products are Python."""

    def __init__(self, ref, n):
        self.ref, self.n, self.rel = ref, n, {}
        self.d = ref.ref_dsf_create()
        for i in range(n):
            ref.ref_dsf_add_element(self.d, i)

    def set_pose(self, m, n, T):
        self.rel[(m, n)] = T.copy()
        self.ref.ref_dsf_union_sets(self.d, max(m, n), min(m, n))          # :167

    def is_exist(self, m, n):
        if m < 0 or n < 0: return False
        if m == n: return True
        if m >= self.n or n >= self.n: return False
        return self.ref.ref_dsf_find_set(self.d, m) == self.ref.ref_dsf_find_set(self.d, n)

    def get_pose(self, m, n):
        if m == n: return np.eye(4)
        assert self.is_exist(m, n)
        if (m, n) in self.rel: return self.rel[(m, n)]
        if (n, m) in self.rel: return np.linalg.inv(self.rel[(n, m)])
        setid = self.ref.ref_dsf_find_set(self.d, m)
        G = self.ref.ref_graph_create(self.n)                               # :69-84: std::map key order, both directions
        for (a, b) in sorted(self.rel):
            if self.ref.ref_dsf_find_set(self.d, a) == setid and self.ref.ref_dsf_find_set(self.d, b) == setid:
                self.ref.ref_graph_add_edge(G, a, b); self.ref.ref_graph_add_edge(G, b, a)
        self.ref.ref_graph_bfs(G, n)                                        # :89
        buf = (C.c_int * 128)(); k = self.ref.ref_graph_get_path_from(G, m, buf, 128); path = list(buf[:k])
        self.ref.ref_graph_destroy(G)
        ans = np.eye(4)
        for h in range(len(path) - 1):                                      # :103-126
            key = (path[h], path[h + 1])
            ans = ans @ (self.rel[key] if key in self.rel else np.linalg.inv(self.rel[(key[1], key[0])]))
        self.set_pose(path[0], path[-1], ans)                               # :137, memoised
        return ans


@pytest.mark.parametrize("seed", range(5))
def test_inferred_world_poses_follow_the_reference_bfs_path(ref, ours, seed):
    """Random relative poses between worlds, deliberately INCONSISTENT around cycles, so that a different path through the
    known pairs gives a visibly different transform: the product's Worlds and the Python front-end must pick the path the
    reference's graph class picks, query after query (answers are memoised, so order matters)."""
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(4, 10))
    model = ReferenceWorldsModel(ref, n)
    W = ours.ours_worlds_create(); P = frontend.Worlds()
    for i in range(n):
        ours.ours_worlds_world_starts(W, 10**9 * (i + 1)); P.world_starts(10**9 * (i + 1))
    T16 = (C.c_double * 16)()
    for _ in range(int(rng.integers(n - 1, 2 * n))):
        m, k = int(rng.integers(0, n)), int(rng.integers(0, n))
        if m == k or (m, k) in model.rel:
            continue
        T = _rand_pose(rng)
        model.set_pose(m, k, T); P.setPoseBetweenWorlds(m, k, T)
        assert ours.ours_worlds_set_pose(W, m, k, np.ascontiguousarray(T).ctypes.data_as(C.POINTER(C.c_double)))
    assert [ours.ours_worlds_find_setid(W, i) for i in range(n)] == [ref.ref_dsf_find_set(model.d, i) for i in range(n)] == [P.find_setID_of_world_i(i) for i in range(n)]
    queries = [(int(a), int(b)) for a, b in rng.integers(0, n, size=(40, 2))]
    inferred = 0
    for m, k in queries:
        assert bool(ours.ours_worlds_is_exist(W, m, k)) == model.is_exist(m, k) == P.is_exist(m, k)
        if not model.is_exist(m, k):
            continue
        inferred += m != k and (m, k) not in model.rel and (k, m) not in model.rel
        want = model.get_pose(m, k)
        assert ours.ours_worlds_get_pose(W, m, k, T16) == 1
        assert np.allclose(np.array(T16[:]).reshape(4, 4), want, rtol=0, atol=1e-9 * max(1.0, np.abs(want).max())), (m, k)
        assert np.allclose(P.getPoseBetweenWorlds(m, k), want, rtol=0, atol=1e-9 * max(1.0, np.abs(want).max())), (m, k)
    assert ours.ours_worlds_n_keys(W) == len(model.rel) == len(P.rel)
    assert inferred > 0, "no query needed the BFS branch: the test would not discriminate"
    ours.ours_worlds_destroy(W); ref.ref_dsf_destroy(model.d)
