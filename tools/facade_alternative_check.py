"""GPU check that has not been run yet (round 1 ended without GPU time): the facade-level evaluation of the reference's
switched-off functors (pgs_facade_evaluate_alternative) against the oracle, after a device solve of a four-world session.
Run on a B200 box:  python tools/facade_alternative_check.py   -> prints one line per functor kind, exits non-zero on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend, pgo  # noqa: E402
from solve_keyframe_pose_graph_b200 import facade, synth  # noqa: E402

g = synth.generate_config(4, n_nodes=120, n_interworld=18)
F = facade.Facade(odom_fanout=3); F.ingest(g)
M = frontend.Manager(); M.ingest(g)
R = frontend.ReferenceFrontEnd(M, odom_fanout=3)
assert F.solve_once() and R.trigger(solve=True) is not None
bad = 0
for kind in (0, 1, 2):
    terms = F.alternative_terms(kind)
    want = pgo.fourdof_eval(kind, **terms)                       # oracle on the facade's own blocks: isolates the kernel call
    got = F.evaluate_alternative(kind)
    ref = pgo.fourdof_eval(kind, **R.alternative_terms(kind))    # the front-end restatement's own session
    dr = np.abs(got["r"] - want["r"]).max() / max(1.0, np.abs(want["r"]).max())
    dJ = np.abs(got["J"] - want["J"]).max() / max(1.0, np.abs(want["J"]).max())
    dc = abs(got["cost"] - want["cost"]) / max(1.0, want["cost"])
    ds = abs(got["cost"] - ref["cost"]) / max(1.0, ref["cost"])
    ok = dr <= 1e-12 and dJ <= 1e-12 and dc <= 1e-12 and ds <= 1e-3
    bad += not ok
    print(f"kind {kind}: blocks {len(terms['c1'])}  r {dr:.2e}  J {dJ:.2e}  cost {dc:.2e}  session cost vs front-end {ds:.2e}  {'ok' if ok else 'MISMATCH'}")
F.close()
sys.exit(1 if bad else 0)
