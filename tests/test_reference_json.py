"""The product's state files against THE JSON LIBRARY THE REFERENCE WRITES THEM WITH.

oracle/_ref/libref_json.so is the reference's vendored nlohmann/json.hpp (3.4.0) compiled where it lies under
/root/reference (recipe: oracle/Makefile, wrapper oracle/ref_json_capi.cpp).  Every file the product's writer
(csrc/host/Json.h + GraphIO.cpp) produces is parsed by that library and dumped again with the reference's `dump(4)`:
byte-identical text means the reference's build would have written exactly these bytes for the same values (key order,
indentation, integer / double formatting) and that its loader accepts them.  Built only where the reference tree exists;
the tests skip without the library."""
import ctypes as C
import os

import numpy as np
import pytest

from solve_keyframe_pose_graph_b200 import facade, synth

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_json.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_json.so not built (needs /root/reference; make -C oracle)")


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(REF_SO)
    L.ref_json_redump.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.ref_json_dump_double.argtypes = [C.c_double, C.c_char_p, C.c_int]
    L.ref_json_dump_int64.argtypes = [C.c_longlong, C.c_char_p, C.c_int]
    L.ref_json_parse_double.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
    return L


def redump(L, text, indent=4):
    n = L.ref_json_redump(text, indent, None, 0)
    if n < 0:
        return None
    buf = C.create_string_buffer(n + 1)
    L.ref_json_redump(text, indent, buf, n + 1)
    return buf.value


def test_state_files_are_what_the_reference_library_would_have_written(ref, tmp_path):
    g = synth.generate_config(4, n_nodes=40, n_interworld=9)
    F = facade.Facade(odom_fanout=3, dry_run=True); F.ingest(g); assert F.solve_once()
    rng = np.random.default_rng(5)
    F.camera_pose_callback(int(g["stamps"][-1]) + 10**8, [1.5, -2.25, 1e-7], [0, 0, 0, 1.0], rng.normal(size=(6, 6)) * 1e-3)   # a covariance with awkward doubles
    F.save_json(tmp_path); F.close()
    for name in ("log_posegraph.json", "log_optimized_poses.json", "solved_posegraph.json"):
        ours = open(tmp_path / name, "rb").read()
        again = redump(ref, ours)
        assert again is not None, f"{name}: the reference's JSON library rejects the file"
        assert again == ours.rstrip(b"\n"), f"{name}: differs from nlohmann dump(4) of the same values"


def test_the_sample_quoted_in_the_reference_source_survives_both_libraries(ref, tmp_path):
    """tests/golden/reference_solved_posegraph_sample.json (from src/NodeDataManager.cpp:892-995): the product's reader
    and writer and the reference's library agree on it value for value."""
    sample = open(os.path.join(HERE, "golden", "reference_solved_posegraph_sample.json"), "rb").read()
    canon = redump(ref, sample)
    assert canon is not None
    T, st, w, sid = facade.io_load_solved_posegraph(os.path.join(HERE, "golden", "reference_solved_posegraph_sample.json"))
    assert len(T) > 0 and np.isfinite(T).all()
    p = tmp_path / "canon.json"; p.write_bytes(canon)
    T2, st2, w2, sid2 = facade.io_load_solved_posegraph(p)
    assert np.array_equal(T, T2) and np.array_equal(st, st2) and np.array_equal(w, w2) and np.array_equal(sid, sid2)


def test_numbers_print_like_the_reference_library(ref, tmp_path):
    """Doubles: nlohmann prints the shortest text that round-trips, with its own rules for exponents, "-0.0" and integers
    stored as doubles.  Awkward values are routed through the product's writer as loop-edge weights and ROS-epoch time
    stamps (both plain JSON numbers in log_posegraph.json) and the file must still be byte-identical to the library's dump."""
    rng = np.random.default_rng(11)
    vals = [0.0, -0.0, 1.0, -1.0, 0.1, 1e-7, 1.5e-5, 123456.789, 1e15, 1e16, 1.7976931348623157e308, 5e-324, 2.5e-3, 1e21, 1e-5, 0.001, 99999999999999.98,
            1523613562.8960001, 3.0e10, 1 / 3, 2 / 3, 1e22, 1e23, 4.35, 0.3] + list(rng.normal(size=150) * 10.0 ** rng.integers(-12, 12, size=150))
    n = len(vals) + 1
    F = facade.Facade(dry_run=True)
    stamps = 1523613562 * 10**9 + np.cumsum(rng.integers(1, 10**9, size=n)).astype(np.int64)          # seconds.nanoseconds as doubles in the file
    q = np.tile([0, 0, 0, 1.0], (n, 1)); t = rng.normal(size=(n, 3))
    F.add_nodes(stamps, q, t)
    F.add_loop_edges(np.arange(1, n), np.zeros(n - 1, np.int32), q[1:], t[1:], np.array(vals))
    F.save_json(tmp_path); F.close()
    ours = open(tmp_path / "log_posegraph.json", "rb").read()
    again = redump(ref, ours)
    assert again is not None and again == ours.rstrip(b"\n")
    import json
    J = json.loads(ours)
    assert [e["weight"] for e in J["loopedges"]] == [float(v) for v in vals]                          # and every value survived exactly
    buf = C.create_string_buffer(64)
    for v in (0, 1, -1, 2**31, -2**31, 2**53 + 1, 1523613562896000100, -9223372036854775807):
        ref.ref_json_dump_int64(v, buf, 64); assert buf.value == str(v).encode()


def test_double_printer_equals_the_reference_library_on_random_bit_patterns(ref, tmp_path):
    """csrc/host/Grisu2.h against nlohmann 3.4.0's own output: 60 000 random IEEE-754 bit patterns over the whole range
    (denormals included), both signs, plus the known hard cases.  (The same comparison ran over 4 million values.)"""
    import struct
    import subprocess
    so = str(tmp_path / "grisu2_hostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "grisu2_hostcheck.cpp")])
    O = C.CDLL(so)
    rng = np.random.default_rng(23)
    bits = rng.integers(0, 2**63, size=30000, dtype=np.uint64)
    vals = np.array([struct.unpack("<d", struct.pack("<Q", int(b)))[0] for b in bits])
    vals = vals[np.isfinite(vals)]
    hard = np.array([1e23, 3e10, 1e-4, 1e-5, 123456789012345678.0, 5e-324, 2.2250738585072014e-308, 2.225073858507201e-308, 1.7976931348623157e308, 1e15, 1e16,
                     9007199254740993.0, 0.3, 4.35, 1 / 3, 1523613562.8960001, 0.1, 2.5e-3])
    vals = np.ascontiguousarray(np.concatenate([hard, -hard, vals, -vals]))
    out = C.create_string_buffer(40 * len(vals))
    assert O.ours_dump_doubles(C.c_int(len(vals)), vals.ctypes.data_as(C.POINTER(C.c_double)), out, C.c_int(40)) == 0
    buf = C.create_string_buffer(64)
    for i, v in enumerate(vals):
        ref.ref_json_dump_double(float(v), buf, 64)
        ours = out.raw[40 * i:40 * i + 40].split(b"\0", 1)[0]
        assert ours == buf.value, (float(v), ours, buf.value)
        assert float(ours) == float(v)                                       # and it round-trips
    assert out.raw[:40].split(b"\0", 1)[0] == b"9.999999999999999e+22"       # Grisu2 is not shortest for 1e23: neither are we
