"""Host-side logic of the multi-GPU path on CPU: the partition rule (libpgs' host-only pgs_partition) and a
world-size-2 gloo run of the border-Schur algebra against the oracle's full-system LM step."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util_graphs import random_graph

import solve_keyframe_pose_graph_b200 as pgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_rule_properties(world):
    g = random_graph(400, 3, 120, seed=21)
    N = g["N"]
    P = pgs.partition(N, world, g["oc1"], g["oc2"], g["la"], g["lb"], g["rn"])
    own = P["node_owner"]
    cut = list(P["cut"])
    assert cut[0] == 0 and cut[-1] == N and all(a < b for a, b in zip(cut, cut[1:]))
    if world >= 3:
        # ranges between the two free ends carry a separator through their elimination: 1 / 3.0 of the nodes (kCarryCost)
        sizes = np.diff(cut)
        assert abs(sizes[0] - sizes[-1]) <= 1 and all(abs(3.0 * m - sizes[0]) <= 3 for m in sizes[1:-1])
        assert list(P["chain_down"]) == [0] * (world - 1) + [1]      # the last range burns downwards from the free top end
    rng = np.searchsorted(cut, np.arange(N), side="right") - 1
    # interior nodes stay in their range; removing the border disconnects the ranges
    assert np.array_equal(own[own >= 0], rng[own >= 0])
    assert P["n_border"] == int((own < 0).sum())
    if world == 1:
        assert P["n_border"] == 0
    pairs = list(zip(g["oc1"], g["oc2"], P["odom_owner"])) + list(zip(g["la"], g["lb"], P["loop_owner"]))
    for i, j, o in pairs:
        assert 0 <= o < world
        if own[i] >= 0 and own[j] >= 0:
            assert own[i] == own[j] == o              # no edge between two different interiors
        for v in (i, j):
            assert own[v] in (-1, o)                  # endpoints are the owner's interior or border
    # a node is border exactly when it has a neighbour in a lower range
    has_lower = np.zeros(N, bool)
    for i, j, _ in pairs:
        if rng[i] > rng[j]:
            has_lower[i] = True
        if rng[j] > rng[i]:
            has_lower[j] = True
    assert np.array_equal(own < 0, has_lower)
    assert np.array_equal(P["reg_owner"], rng[g["rn"]])


@pytest.mark.parametrize("chains", [2, 4])
def test_single_gpu_chain_plan(chains):
    """One GPU, several chains: the same rule over `chains` ranges that all belong to rank 0."""
    g = random_graph(600, 3, 150, seed=5)
    N = g["N"]
    P = pgs.partition(N, 1, g["oc1"], g["oc2"], g["la"], g["lb"], g["rn"], chains_per_rank=chains)
    assert P["n_chains"] == chains and list(P["cut"]) == [0, N]
    ch = P["node_chain"]
    assert set(ch[ch >= 0]) == set(range(chains)) and P["n_border"] == int((ch < 0).sum()) > 0
    assert list(P["chain_down"]) == [0] * (chains - 1) + [1]
    # interiors of different chains are never adjacent
    for i, j in list(zip(g["oc1"], g["oc2"])) + list(zip(g["la"], g["lb"])):
        if ch[i] >= 0 and ch[j] >= 0:
            assert ch[i] == ch[j]
    assert (P["node_owner"][ch >= 0] == 0).all() and (P["odom_owner"] == 0).all()
    if chains == 2:     # burn at both ends: two halves, nothing carried
        first = np.nonzero(ch == 1)[0].min()
        assert abs(first - N // 2) < 60


def test_partition_rejects_bad_indices():
    with pytest.raises(pgs.PgsError):
        pgs.partition(10, 2, [0, 11], [1, 2], [], [], [])


@pytest.mark.parametrize("world,port", [(2, 29533), (3, 29534)])
def test_border_schur_scheme_gloo(world, port):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("dist-cpu seed") == 2, r.stdout
