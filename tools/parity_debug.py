#!/usr/bin/env python
"""Side-by-side LM tables of the CUDA solve and the oracle on one seeded graph (debugging aid for parity mismatches)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util_graphs import load_oracle, load_pgs, random_graph, rot_angle_between

n, fan, nl, of, seed = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (700, 3, 120, 0.1, 1)
g = random_graph(n, fan, nl, outlier_frac=of, seed=seed)
O = load_oracle(g); so = O.solve(); qo, to = O.poses()
S = load_pgs(g, chains=1); ss = S.solve(); qs, ts = S.poses()
print("it | oracle cost rho radius step ok | gpu cost rho radius step ok")
for a, b in zip(so["iterations"], ss["iterations"]):
    print(f"{a['iteration']:2d} | {a['cost']:.15g} {a['relative_decrease']:.6g} {a['trust_region_radius']:.6g} {a['step_norm']:.8g} {a['step_is_successful']} | "
          f"{b['cost']:.15g} {b['relative_decrease']:.6g} {b['trust_region_radius']:.6g} {b['step_norm']:.8g} {b['step_is_successful']}")
print("dt", np.abs(ts - to).max(), "drot", rot_angle_between(qs, qo).max(), "backward errors", S.linear_backward_errors())
# the first LM step from the same point, three ways
S2 = load_pgs(g, chains=1); O2 = load_oracle(g)
for radius in (1e4, 1e7):
    dp, ds, mcc, _ = S2.linear_step(radius)
    dpo, dso, mcco = O2.linear_step(radius)
    print("radius", radius, "step rel diff", np.abs(dp - dpo).max() / np.abs(dpo).max(), "mcc", mcc, mcco)
