"""ctypes front-end of the CPU ORACLE (oracle/sgz_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg -- never by strugatzki_b200/.
PARITY UNPINNED: see oracle/sgz_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsgz_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sgz_oracle.c")
    hdr = os.path.join(_HERE, "sgz_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        if not os.path.exists(src):
            raise RuntimeError("oracle sources missing and no prebuilt libsgz_oracle.so")
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsgz_oracle.so"])
    return _LIB_PATH


class _Match(C.Structure):
    _fields_ = [("sim", C.c_float), ("file", C.c_int32), ("start", C.c_int64), ("stop", C.c_int64),
                ("boostIn", C.c_float), ("boostOut", C.c_float)]


class _Break(C.Structure):
    _fields_ = [("sim", C.c_float), ("_pad", C.c_int32), ("pos", C.c_int64)]


class _CorrCfg(C.Structure):
    _fields_ = [("numCh", C.c_int32), ("stepSize", C.c_int32), ("norm", C.c_void_p),
                ("input", C.c_void_p), ("inputFrames", C.c_int64),
                ("punchInStart", C.c_int64), ("punchInStop", C.c_int64), ("punchInWeight", C.c_float),
                ("hasPunchOut", C.c_int32),
                ("punchOutStart", C.c_int64), ("punchOutStop", C.c_int64), ("punchOutWeight", C.c_float),
                ("minPunch", C.c_int64), ("maxPunch", C.c_int64),
                ("maxBoost", C.c_float), ("numMatches", C.c_int32), ("numPerFile", C.c_int32),
                ("minSpacing", C.c_int64)]


class _SegmCfg(C.Structure):
    _fields_ = [("numCh", C.c_int32), ("stepSize", C.c_int32), ("norm", C.c_void_p),
                ("hasStart", C.c_int32), ("hasStop", C.c_int32),
                ("spanStart", C.c_int64), ("spanStop", C.c_int64), ("corrLen", C.c_int64),
                ("temporalWeight", C.c_float), ("numBreaks", C.c_int32), ("minSpacing", C.c_int64)]


class _SelfCfg(C.Structure):
    _fields_ = [("numCh", C.c_int32), ("stepSize", C.c_int32), ("norm", C.c_void_p),
                ("hasStart", C.c_int32), ("hasStop", C.c_int32),
                ("spanStart", C.c_int64), ("spanStop", C.c_int64), ("corrLen", C.c_int64),
                ("decimation", C.c_int32), ("temporalWeight", C.c_float), ("colorInv", C.c_int32),
                ("colorWarp", C.c_float), ("colorCeil", C.c_float)]


class _CrossCfg(C.Structure):
    _fields_ = [("numCh", C.c_int32), ("stepSize", C.c_int32), ("norm", C.c_void_p),
                ("has1Start", C.c_int32), ("has1Stop", C.c_int32),
                ("span1Start", C.c_int64), ("span1Stop", C.c_int64),
                ("has2Start", C.c_int32), ("has2Stop", C.c_int32),
                ("span2Start", C.c_int64), ("span2Stop", C.c_int64),
                ("temporalWeight", C.c_float), ("maxBoost", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.sgz_o_corr_search.restype = C.c_int
        _lib.sgz_o_corr_curve.restype = C.c_int64
        _lib.sgz_o_corr_num_offsets.restype = C.c_int64
        _lib.sgz_o_segm_run.restype = C.c_int
        _lib.sgz_o_self_geometry.restype = C.c_int
        _lib.sgz_o_self_image.restype = C.c_int
        _lib.sgz_o_self_cells.restype = C.c_int
        _lib.sgz_o_cross_run.restype = C.c_int64
        _lib.sgz_o_avg.restype = C.c_float
        _lib.sgz_o_correlate_half.restype = C.c_float
        _lib.sgz_o_correlate.restype = C.c_float
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class CorrParams:
    """Mirror of FeatureCorrelation.Config in oracle terms (all spans in sample frames)."""
    step_size: int
    input: np.ndarray                       # [frames][numCh] raw features of metaInput
    punch_in: Tuple[int, int]
    punch_in_weight: float = 0.5
    punch_out: Optional[Tuple[int, int]] = None
    punch_out_weight: float = 0.5
    min_punch: int = 22050
    max_punch: int = 88200
    norm: Optional[np.ndarray] = None       # [numCh][2] or None (normalize=false)
    max_boost: float = 8.0
    num_matches: int = 1
    num_per_file: int = 1
    min_spacing: int = 0
    _keep: list = field(default_factory=list, repr=False)

    def c_struct(self) -> _CorrCfg:
        inp = _f32(self.input)
        nrm = None if self.norm is None else _f32(self.norm)
        self._keep = [inp, nrm]
        po = self.punch_out or (0, 0)
        return _CorrCfg(inp.shape[1], self.step_size, _ptr(nrm), _ptr(inp), inp.shape[0],
                        self.punch_in[0], self.punch_in[1], self.punch_in_weight,
                        0 if self.punch_out is None else 1, po[0], po[1], self.punch_out_weight,
                        self.min_punch, self.max_punch, self.max_boost, self.num_matches,
                        self.num_per_file, self.min_spacing)


def corr_search(p: CorrParams, files: Sequence[np.ndarray]) -> List[dict]:
    """FeatureCorrelationImpl.body() over an ordered in-memory DB; files[i] = [frames][numCh]."""
    cfg = p.c_struct()
    fs = [_f32(f) for f in files]
    n = len(fs)
    ptrs = (C.c_void_p * max(n, 1))(*[f.ctypes.data for f in fs])
    lens = (C.c_int64 * max(n, 1))(*[f.shape[0] for f in fs])
    cap = max(p.num_matches, 1) + 8
    out = (_Match * cap)()
    k = lib().sgz_o_corr_search(C.byref(cfg), n, ptrs, lens, out, cap)
    if k < 0:
        raise RuntimeError(f"oracle corr_search failed: {k}")
    return [dict(sim=out[i].sim, file=out[i].file, start=out[i].start, stop=out[i].stop,
                 boostIn=out[i].boostIn, boostOut=out[i].boostOut) for i in range(k)]


def corr_curve(p: CorrParams, file: np.ndarray, which: int = 0, first_frame: int = 0):
    cfg = p.c_struct()
    f = _f32(file)
    cap = max(f.shape[0], 1)
    sim = np.empty(cap, np.float32)
    boost = np.empty(cap, np.float32)
    k = lib().sgz_o_corr_curve(C.byref(cfg), which, _ptr(f), C.c_int64(f.shape[0]), C.c_int64(first_frame),
                               _ptr(sim), _ptr(boost), C.c_int64(cap))
    if k < 0:
        raise RuntimeError(f"oracle corr_curve failed: {k}")
    return sim[:k].copy(), boost[:k].copy()


def corr_num_offsets(p: CorrParams, n_frames: Sequence[int]) -> int:
    cfg = p.c_struct()
    lens = (C.c_int64 * max(len(n_frames), 1))(*n_frames)
    return int(lib().sgz_o_corr_num_offsets(C.byref(cfg), len(n_frames), lens))


@dataclass
class SegmParams:
    step_size: int
    corr_len: int = 22050
    temporal_weight: float = 0.5
    norm: Optional[np.ndarray] = None
    num_breaks: int = 1
    min_spacing: int = 22050
    span_start: Optional[int] = None
    span_stop: Optional[int] = None


def segm_run(p: SegmParams, file: np.ndarray, want_curve: bool = False):
    f = _f32(file)
    nrm = None if p.norm is None else _f32(p.norm)
    cfg = _SegmCfg(f.shape[1], p.step_size, _ptr(nrm), int(p.span_start is not None),
                   int(p.span_stop is not None), p.span_start or 0, p.span_stop or 0, p.corr_len,
                   p.temporal_weight, p.num_breaks, p.min_spacing)
    cap = max(p.num_breaks, 1) + 8
    out = (_Break * cap)()
    curve = np.full(max(f.shape[0], 1), np.nan, np.float32) if want_curve else None
    k = lib().sgz_o_segm_run(C.byref(cfg), _ptr(f), C.c_int64(f.shape[0]), out, cap, _ptr(curve),
                             C.c_int64(0 if curve is None else curve.shape[0]))
    if k < 0:
        raise RuntimeError(f"oracle segm_run failed: {k}")
    breaks = [dict(sim=out[i].sim, pos=out[i].pos) for i in range(k)]
    return (breaks, curve) if want_curve else breaks


@dataclass
class SelfParams:
    step_size: int
    corr_len: int = 44100
    decimation: int = 1
    temporal_weight: float = 0.5
    norm: Optional[np.ndarray] = None
    color_inv: bool = False
    color_warp: float = 1.0
    color_ceil: float = 1.0
    span_start: Optional[int] = None
    span_stop: Optional[int] = None


def _self_cfg(p: SelfParams, num_ch: int, keep: list) -> _SelfCfg:
    nrm = None if p.norm is None else _f32(p.norm)
    keep.append(nrm)
    return _SelfCfg(num_ch, p.step_size, _ptr(nrm), int(p.span_start is not None),
                    int(p.span_stop is not None), p.span_start or 0, p.span_stop or 0, p.corr_len,
                    p.decimation, p.temporal_weight, int(p.color_inv), p.color_warp, p.color_ceil)


def self_geometry(p: SelfParams, num_ch: int, n1: int, n2: int):
    keep: list = []
    cfg = _self_cfg(p, num_ch, keep)
    d, nc, st = C.c_int32(), C.c_int32(), C.c_int32()
    ext = lib().sgz_o_self_geometry(C.byref(cfg), C.c_int64(n1), C.c_int64(n2), C.byref(d), C.byref(nc),
                                    C.byref(st))
    if ext < 0:
        raise RuntimeError(f"oracle self_geometry failed: {ext}")
    return dict(imgExt=ext, decim=d.value, numCorrs=nc.value, afStart=st.value)


def self_image(p: SelfParams, file1: np.ndarray, file2: Optional[np.ndarray] = None) -> np.ndarray:
    f1 = _f32(file1)
    f2 = f1 if file2 is None else _f32(file2)
    keep: list = []
    cfg = _self_cfg(p, f1.shape[1], keep)
    g = self_geometry(p, f1.shape[1], f1.shape[0], f2.shape[0])
    ext = g["imgExt"]
    rgb = np.zeros((max(ext, 1), max(ext, 1)), np.int32)
    k = lib().sgz_o_self_image(C.byref(cfg), _ptr(f1), C.c_int64(f1.shape[0]), _ptr(f2),
                               C.c_int64(f2.shape[0]), _ptr(rgb), C.c_int64(rgb.size))
    if k < 0:
        raise RuntimeError(f"oracle self_image failed: {k}")
    return rgb[:ext, :ext]


def self_cells(p: SelfParams, file1: np.ndarray, file2: Optional[np.ndarray], left: np.ndarray,
               right: np.ndarray):
    f1 = _f32(file1)
    f2 = f1 if file2 is None else _f32(file2)
    keep: list = []
    cfg = _self_cfg(p, f1.shape[1], keep)
    l = np.ascontiguousarray(left, np.int32)
    r = np.ascontiguousarray(right, np.int32)
    sim = np.empty(l.shape[0], np.float32)
    rgb = np.empty(l.shape[0], np.int32)
    k = lib().sgz_o_self_cells(C.byref(cfg), _ptr(f1), C.c_int64(f1.shape[0]), _ptr(f2),
                               C.c_int64(f2.shape[0]), C.c_int64(l.shape[0]), _ptr(l), _ptr(r),
                               _ptr(sim), _ptr(rgb))
    if k < 0:
        raise RuntimeError(f"oracle self_cells failed: {k}")
    return sim, rgb


@dataclass
class CrossParams:
    step_size: int
    temporal_weight: float = 0.5
    norm: Optional[np.ndarray] = None
    max_boost: float = 8.0
    span1: tuple = (None, None)      # (start, stop) in sample frames; None = open end (Span.All / HasStart / HasStop)
    span2: tuple = (None, None)


def cross_run(p: CrossParams, file1: np.ndarray, file2: np.ndarray) -> np.ndarray:
    """CrossSimilarityImpl.body(): the sim curve the reference writes to its 1-channel output file."""
    f1, f2 = _f32(file1), _f32(file2)
    nrm = None if p.norm is None else _f32(p.norm)
    (a1, b1), (a2, b2) = p.span1, p.span2
    cfg = _CrossCfg(f1.shape[1], p.step_size, _ptr(nrm), int(a1 is not None), int(b1 is not None), a1 or 0, b1 or 0,
                    int(a2 is not None), int(b2 is not None), a2 or 0, b2 or 0, p.temporal_weight, p.max_boost)
    cap = max(f1.shape[0], f2.shape[0], 1) + 1
    out = np.zeros(cap, np.float32)
    n = lib().sgz_o_cross_run(C.byref(cfg), _ptr(f1), C.c_int64(f1.shape[0]), _ptr(f2), C.c_int64(f2.shape[0]),
                              _ptr(out), C.c_int64(cap))
    if n < 0:
        raise RuntimeError(f"oracle cross_run failed: {n}")
    return out[:n].copy()


def stats_run(files: Sequence[np.ndarray], want_per_file: bool = False):
    """FeatureStatsImpl.body(): [numCh][2] doubles = (min of the files' 1st percentiles, max of their 99th)."""
    fs = [_f32(f) for f in files]
    num_ch = fs[0].shape[1]
    ptrs = (C.c_void_p * len(fs))(*[f.ctypes.data for f in fs])
    nfr = (C.c_int64 * len(fs))(*[f.shape[0] for f in fs])
    out = np.zeros((num_ch, 2), np.float64)
    per = np.zeros((len(fs), num_ch, 2), np.float64) if want_per_file else None
    lib().sgz_o_stats_run(num_ch, len(fs), ptrs, nfr, out.ctypes.data_as(C.c_void_p),
                          None if per is None else per.ctypes.data_as(C.c_void_p))
    return (out, per) if want_per_file else out
