"""Times the alternative-functor kernel (include/pgs_fourdof.h) on a c3-sized edge list; prints one JSON line per kind."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import solve_keyframe_pose_graph_b200.capi as capi  # noqa: E402
from test_fourdof import random_edges  # noqa: E402

for kind, name in ((0, "FourDOFError"), (1, "FourDOFErrorWithSwitchingConstraints"), (2, "QinFourDOFWeightError")):
    NR, NC, RW = capi.FOURDOF_SHAPES[kind]
    g = random_edges(kind, 100000, 350000, seed=kind)
    ms = min(capi.fourdof_evaluate(kind, **g)["ms_kernel"] for _ in range(5))
    per_edge = 8 + 8 * RW + 24 + 8 + 8 * NR + 8 * NR * NC          # indices, observation, weight/switch, r, J
    algo = 350000 * per_edge + 100000 * 8 * (RW + 3)
    print(json.dumps(dict(functor=name, edges=350000, ms_kernel=ms, edge_evals_per_s=350000 / ms * 1e3, algorithmic_GBs=algo / ms / 1e6, bytes_per_edge=per_edge)))
