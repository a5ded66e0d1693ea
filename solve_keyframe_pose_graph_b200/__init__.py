"""B200-native pose-graph hot path behind the NodeDataManager / PoseGraphSLAM API surface of
mpkuse/solve_keyframe_pose_graph.  Python here is a thin ctypes view of libpgs.so (C-ABI in
include/pgs.h); all compute is hand-written sm_100a CUDA.  There is no CPU fallback."""
from .capi import (Options, PoseGraphSolver, PgsError, Summary, Iteration, lib, library_path,  # noqa: F401
                   exported_symbols, dist_unique_id, partition)
