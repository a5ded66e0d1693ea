"""ctypes view of include/pgs_synth.h — the deterministic graph generator of the bench harness."""
import ctypes as C

import numpy as np

from .capi import lib


class SynthSpec(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("n_loop", C.c_int32), ("n_worlds", C.c_int32), ("n_interworld", C.c_int32),
        ("deadzone_nodes", C.c_int32), ("loop_gap_min", C.c_int32), ("loop_gap_max", C.c_int32),
        ("outlier_fraction", C.c_double), ("odom_sigma_t", C.c_double), ("odom_sigma_r", C.c_double),
        ("loop_sigma_t", C.c_double), ("loop_sigma_r", C.c_double), ("seed", C.c_uint64),
    ]


def config_spec(config, **overrides):
    s = SynthSpec()
    if lib().pgs_synth_config(C.c_int32(config), C.byref(s)) != 0:
        raise ValueError(f"unknown config {config}")
    for k, v in overrides.items():
        setattr(s, k, v)
    return s


def generate(spec):
    """Returns a dict of numpy arrays: stamps, q, t (manager poses), gt_q, gt_t, la, lb, lq, lt, lw, lout, k0, k1."""
    L = lib()
    h = C.c_void_p()
    if L.pgs_synth_create(C.byref(spec), C.byref(h)) != 0:
        raise ValueError("pgs_synth_create failed")
    try:
        n = C.c_int32(); e = C.c_int32(); k = C.c_int32()
        L.pgs_synth_sizes(h, C.byref(n), C.byref(e), C.byref(k))
        N, E, K = n.value, e.value, k.value
        g = dict(N=N, stamps=np.zeros(N, np.int64), q=np.zeros((N, 4)), t=np.zeros((N, 3)), gt_q=np.zeros((N, 4)), gt_t=np.zeros((N, 3)),
                 la=np.zeros(E, np.int32), lb=np.zeros(E, np.int32), lq=np.zeros((E, 4)), lt=np.zeros((E, 3)), lw=np.zeros(E),
                 lout=np.zeros(E, np.uint8), k0=np.zeros(K, np.int64), k1=np.zeros(K, np.int64))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        L.pgs_synth_copy(h, p(g["stamps"]), p(g["q"]), p(g["t"]), p(g["gt_q"]), p(g["gt_t"]), p(g["la"]), p(g["lb"]), p(g["lq"]), p(g["lt"]),
                         p(g["lw"]), p(g["lout"]), p(g["k0"]), p(g["k1"]))
        g["lout"] = g["lout"].astype(bool)
        return g
    finally:
        L.pgs_synth_destroy(h)


def generate_config(config, **overrides):
    return generate(config_spec(config, **overrides))
