"""The C-ABI library loads on a CPU-only box and exports every symbol include/strugatzki_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from util import N, ROOT

HEADER = os.path.join(ROOT, "include", "strugatzki_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgz_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    N.build()
    assert os.path.exists(N.LIB_PATH)
    assert N.lib().sgz_abi_version() == 1


def test_every_declared_symbol_is_exported():
    funcs = declared_functions()
    assert len(funcs) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", N.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [f for f in funcs if f not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(N.SYMBOLS) == funcs, "strugatzki_b200/_native.py SYMBOLS out of sync with the header"


def test_struct_layouts_match_header():
    # sizes the Scala/JNI or ctypes side relies on
    assert C.sizeof(N.Match) == 32 and C.sizeof(N.Break) == 16 and C.sizeof(N.Record) == 32
    assert C.sizeof(N.FileSummary) == 16 and C.sizeof(N.SelfGeometry) == 24
    assert C.sizeof(N.CorrConfig) == 96 and C.sizeof(N.SegmConfig) == 56 and C.sizeof(N.CrossConfig) == 64


def test_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", N.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    rc = N.lib().sgz_ctx_create(0, C.byref(ctx))
    assert rc == N.ERR_CUDA
    assert b"no CPU fallback" in N.lib().sgz_last_error()
    from strugatzki_b200 import NativeError, engine
    with pytest.raises(NativeError):
        engine.Context(0)


def test_geometry_needs_no_gpu():
    # pure host helper of the ABI: SelfSimilarityImpl.scala:75-91 incl. the 0xB504 auto-decimation
    from strugatzki_b200 import engine
    cfg = N.SelfConfig(512, 0, 0, 0, 0, 44100, 1, 0.5, 0, 1.0, 1.0, None, 0)
    g = engine.self_geometry(cfg, 155000, 155000)
    assert g == dict(imgExt=38707, decim=4, numCorrs=154829, afStart=0, numCells=38707 * 38708 // 2)
    g = engine.self_geometry(cfg, 1000, 900)
    assert (g["imgExt"], g["decim"], g["numCorrs"]) == (729, 1, 729)


def test_cross_output_count_needs_no_gpu():
    # CrossSimilarityImpl.scala:127-171: the first read swallows min(len2, 8192) frames, then one value per frame
    n = C.c_int64()
    cfg = N.CrossConfig(512, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.5, 8.0)
    for n1, n2, want in ((300, 9000, 809), (9000, 300, 809), (300, 8192, 1), (300, 500, 1), (0, 0, 0)):
        assert N.lib().sgz_cross_num_outputs(C.byref(cfg), C.c_int64(n1), C.c_int64(n2), C.byref(n)) == 0
        assert n.value == want, (n1, n2, n.value)
    # spans: [100*512, 400*512) of file 1 -> 300 frames
    cfg = N.CrossConfig(512, 1, 1, 0, 0, 0, 100 * 512, 400 * 512, 0, 0, 0.5, 8.0)
    assert N.lib().sgz_cross_num_outputs(C.byref(cfg), C.c_int64(5000), C.c_int64(10000), C.byref(n)) == 0
    assert n.value == 1 + 10000 - 8192
    cfg = N.CrossConfig(512, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.5, 8.0)
    assert N.lib().sgz_cross_num_outputs(C.byref(cfg), C.c_int64(9000), C.c_int64(10000), C.byref(n)) == N.ERR_INVALID
    assert b"8192" in N.lib().sgz_last_error()
    assert N.lib().sgz_cross_num_outputs(C.byref(cfg), C.c_int64(0), C.c_int64(100), C.byref(n)) == N.ERR_INVALID
