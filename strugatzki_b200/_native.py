"""ctypes binding of the C ABI in include/strugatzki_b200.h (libsgz_b200.so).

The library is the product; there is no Python or CPU fallback.  Importing this module never
touches the GPU; the first compute call fails loudly (NativeError) when no B200 is visible or
when the shared library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgz_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_ABORTED, ERR_STATE, ERR_IO = 0, -1, -2, -3, -4, -5, -6
LAYOUT_INTERLEAVED_LE, LAYOUT_INTERLEAVED_BE, LAYOUT_PLANAR_LE = 0, 1, 2
LAYOUT_HOST_STABLE = 0x100   # OR-ed in: host buffer stays valid until finalize (pipelined upload)

# every symbol include/strugatzki_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "sgz_abi_version", "sgz_last_error", "sgz_device_count",
    "sgz_ctx_create", "sgz_ctx_destroy", "sgz_ctx_synchronize", "sgz_ctx_stream", "sgz_ctx_last_timing",
    "sgz_ctx_launch_count", "sgz_self_last_kernel",
    "sgz_ctx_trim",
    "sgz_db_create", "sgz_db_destroy", "sgz_db_reserve", "sgz_db_add_file", "sgz_db_add_file_device",
    "sgz_db_add_synth", "sgz_db_add_synth_many", "sgz_db_patch", "sgz_db_finalize", "sgz_db_finalize_async", "sgz_db_stats", "sgz_db_info", "sgz_db_file_frames", "sgz_db_read",
    "sgz_corr_create", "sgz_corr_destroy", "sgz_corr_run", "sgz_corr_start", "sgz_corr_poll", "sgz_corr_abort",
    "sgz_corr_wait", "sgz_corr_result", "sgz_corr_num_offsets", "sgz_corr_timing", "sgz_corr_curve",
    "sgz_corr_scan", "sgz_corr_local_summary", "sgz_corr_set_global", "sgz_corr_local_top", "sgz_corr_set_global_top", "sgz_corr_local_best", "sgz_corr_finish_from_best", "sgz_corr_select", "sgz_corr_records",
    "sgz_corr_merge",
    "sgz_segm_run", "sgz_self_geometry_of", "sgz_self_run", "sgz_self_cells", "sgz_cross_num_outputs",
    "sgz_cross_run", "sgz_measure_peak",
]


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"strugatzki_b200 native error {code}: {msg}")
        self.code = code


class Aborted(NativeError):
    """Mirror of de.sciss.processor.Processor.Aborted."""


class Match(C.Structure):
    _fields_ = [("sim", C.c_float), ("file", C.c_int32), ("start", C.c_int64), ("stop", C.c_int64),
                ("boostIn", C.c_float), ("boostOut", C.c_float)]


class Break(C.Structure):
    _fields_ = [("sim", C.c_float), ("_pad", C.c_int32), ("pos", C.c_int64)]


class CorrConfig(C.Structure):
    _fields_ = [("stepSize", C.c_int32),
                ("punchInStart", C.c_int64), ("punchInStop", C.c_int64), ("punchInWeight", C.c_float),
                ("hasPunchOut", C.c_int32),
                ("punchOutStart", C.c_int64), ("punchOutStop", C.c_int64), ("punchOutWeight", C.c_float),
                ("minPunch", C.c_int64), ("maxPunch", C.c_int64), ("maxBoost", C.c_float),
                ("numMatches", C.c_int32), ("numPerFile", C.c_int32), ("minSpacing", C.c_int64)]


class SegmConfig(C.Structure):
    _fields_ = [("stepSize", C.c_int32), ("hasStart", C.c_int32), ("hasStop", C.c_int32),
                ("spanStart", C.c_int64), ("spanStop", C.c_int64), ("corrLen", C.c_int64),
                ("temporalWeight", C.c_float), ("numBreaks", C.c_int32), ("minSpacing", C.c_int64)]


class CrossConfig(C.Structure):
    _fields_ = [("stepSize", C.c_int32), ("has1Start", C.c_int32), ("has1Stop", C.c_int32),
                ("has2Start", C.c_int32), ("has2Stop", C.c_int32), ("_pad", C.c_int32),
                ("span1Start", C.c_int64), ("span1Stop", C.c_int64),
                ("span2Start", C.c_int64), ("span2Stop", C.c_int64),
                ("temporalWeight", C.c_float), ("maxBoost", C.c_float)]


class SelfConfig(C.Structure):
    _fields_ = [("stepSize", C.c_int32), ("hasStart", C.c_int32), ("hasStop", C.c_int32),
                ("spanStart", C.c_int64), ("spanStop", C.c_int64), ("corrLen", C.c_int64),
                ("decimation", C.c_int32), ("temporalWeight", C.c_float), ("colorInv", C.c_int32),
                ("colorWarp", C.c_float), ("colorCeil", C.c_float), ("lut", C.c_void_p), ("lutSize", C.c_int32),
                ("precise", C.c_int32)]


class SelfGeometry(C.Structure):
    _fields_ = [("imgExt", C.c_int32), ("decim", C.c_int32), ("numCorrs", C.c_int32), ("afStart", C.c_int32),
                ("numCells", C.c_int64)]


class FileSummary(C.Structure):
    _fields_ = [("maxSim", C.c_float), ("numOffsets", C.c_int32), ("maxSimOut", C.c_float), ("_pad", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("file", C.c_int32), ("kind", C.c_int32), ("piOff", C.c_int32), ("poOff", C.c_int32),
                ("sim", C.c_float), ("boostIn", C.c_float), ("boostOut", C.c_float), ("aux", C.c_int32)]


ENTRY_DTYPE = np.dtype([("file", np.int32), ("maxSim", np.float32)])
SUMMARY_DTYPE = np.dtype([("maxSim", np.float32), ("numOffsets", np.int32), ("maxSimOut", np.float32),
                          ("_pad", np.int32)])
RECORD_DTYPE = np.dtype([("file", np.int32), ("kind", np.int32), ("piOff", np.int32), ("poOff", np.int32),
                         ("sim", np.float32), ("boostIn", np.float32), ("boostOut", np.float32),
                         ("aux", np.int32)])
assert SUMMARY_DTYPE.itemsize == C.sizeof(FileSummary) and RECORD_DTYPE.itemsize == C.sizeof(Record)


def build(force: bool = False) -> str:
    """Compile libsgz_b200.so for sm_100a with nvcc (in-tree; works without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "strugatzki_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SGZ_LIB_PATH", LIB_PATH)   # developer knob: A/B a saved build against the in-tree one
        if not os.path.exists(path):
            raise NativeError(ERR_STATE, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                         f"g.build()'` (there is no CPU fallback)")
        L = C.CDLL(path)
        L.sgz_last_error.restype = C.c_char_p
        L.sgz_ctx_stream.restype = C.c_void_p
        L.sgz_ctx_launch_count.restype = C.c_int64
        L.sgz_ctx_trim.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        for name in SYMBOLS:
            getattr(L, name)  # AttributeError if the header and the library disagree
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().sgz_last_error().decode("utf-8", "replace")
        if rc == ERR_ABORTED:
            raise Aborted(rc, msg or "aborted")
        raise NativeError(rc, msg)
    return rc


def fptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
