"""B200-native pose-graph hot path behind the NodeDataManager / PoseGraphSLAM API surface of
mpkuse/solve_keyframe_pose_graph.  Python here is a thin ctypes view of libpgs.so (C-ABI in
include/pgs.h); all compute is hand-written sm_100a CUDA.  There is no CPU fallback."""
import os as _os

# A factorisation runs on three to seven streams that must overlap (panel chain, next-panel tiles, trailing update, per
# elimination chain).  With the driver's default of 8 hardware work queues per context, streams of a process that also runs
# NCCL end up sharing a queue and serialise each other: measured 56 us per panel instead of 38 on every rank of a 2- and a
# 4-GPU solve (profiles/r02_bench_2gpu_connections.txt).  32 is the driver's maximum; it only takes effect if it is in the
# environment before the CUDA context is created, so libpgs.so sets it when it is loaded as well (csrc/pgs_capi.cu).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .capi import (Options, PoseGraphSolver, PgsError, Summary, Iteration, lib, library_path,  # noqa: F401
                   exported_symbols, dist_unique_id, partition)
