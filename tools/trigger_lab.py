#!/usr/bin/env python
"""Where the wall-clock time of one plugin call goes (host graph -> pgs_facade_solve_once -> poses): the library's own
host laps (PGS_HOST_TIMING=1, stderr) next to the wall clock around the call.  python tools/trigger_lab.py --config 3"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PGS_HOST_TIMING", "1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    from solve_keyframe_pose_graph_b200 import facade, synth
    g = synth.generate_config(args.config)
    for rep in range(args.reps):
        F = facade.Facade(odom_fanout=3, device=0)
        F.ingest(g)
        t0 = time.perf_counter(); ok = F.solve_once(); t1 = time.perf_counter(); F.poses(); t2 = time.perf_counter()
        s = F.summary()
        print(f"rep {rep}: solve_once {1e3 * (t1 - t0):.1f} ms wall, get_poses {1e3 * (t2 - t1):.1f} ms, device LM loop {s['ms_total']:.1f} ms "
              f"({max(1, s['num_iterations'] - 1)} iterations), host share {1e3 * (t2 - t0) - s['ms_total']:.1f} ms", file=sys.stderr)
        F.close()


if __name__ == "__main__":
    main()
