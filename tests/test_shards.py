"""CPU test of the node-range sharding bench.py uses for N > 1: every residual block lands in exactly
one shard, a shard only keeps the poses it touches (own range + halo), and the per-shard costs of the oracle add
up to the cost of the whole graph."""
import numpy as np
import pytest

from oracle import pgo
from solve_keyframe_pose_graph_b200 import problems


def _oracle(p):
    P = pgo.Problem()
    P.set_nodes(p["q"], p["t"])
    if len(p["oc1"]):
        P.add_odom_edges(p["oc1"], p["oc2"], p["oq"], p["ot"], p["ow"])
    if len(p["la"]):
        P.add_loop_edges(p["lb"], p["la"], p["lq"], p["lt"], p["lw"])
    if len(p["rn"]):
        P.set_regularizers(p["rn"], p["rq"], p["rt"], p["rw"])
    return P


@pytest.mark.parametrize("world", [2, 4])
def test_shards_partition_the_blocks_and_the_cost(world):
    pgo.build()
    p = problems.build_problem(2, n_nodes=4000, n_loop=600)
    full = _oracle(p).evaluate(autodiff=False)["cost"]
    tot, n_o, n_l, n_r = 0.0, 0, 0, 0
    for r in range(world):
        s = problems.shard_problem(p, r, world)
        lo, hi = s["shard"][2], s["shard"][3]
        assert s["N"] == len(s["nodes"]) == (hi - lo) + s["n_halo"]
        assert np.array_equal(s["nodes"][: hi - lo], np.arange(lo, hi))                 # own range first, halo above it
        assert s["n_halo"] <= 2000 and (r < world - 1 or s["n_halo"] == 0)             # halo bounded by the loop-gap bound
        for k in ("oc1", "oc2", "la", "lb", "rn"):
            assert len(s[k]) == 0 or (s[k].min() >= 0 and s[k].max() < s["N"])
        assert np.array_equal(s["q"], p["q"][s["nodes"]])
        n_o += len(s["oc1"]); n_l += len(s["la"]); n_r += len(s["rn"])
        tot += _oracle(s).evaluate(autodiff=False)["cost"]
    assert (n_o, n_l, n_r) == (len(p["oc1"]), len(p["la"]), len(p["rn"]))
    assert abs(tot - full) <= 1e-12 * full
