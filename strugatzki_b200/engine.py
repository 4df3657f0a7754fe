"""Thin object wrappers over the C ABI (one class per opaque handle).

Nothing here computes: every method is one call into libsgz_b200.so.  Host arrays are numpy;
device memory never crosses this layer.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N


def _frames(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("feature frames must be a 2-D array [frames][channels]")
    return a


class Context:
    """sgz_ctx: one GPU + one stream."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().sgz_ctx_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            N.lib().sgz_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        N.check(N.lib().sgz_ctx_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(N.lib().sgz_ctx_stream(self._h) or 0)

    def last_timing(self) -> Tuple[float, int]:
        ms, n = C.c_float(), C.c_int64()
        N.check(N.lib().sgz_ctx_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launch_count(self) -> int:
        return int(N.lib().sgz_ctx_launch_count(self._h))

    def trim(self) -> int:
        """Return the device memory parked by destroyed databases / jobs to the driver; bytes freed."""
        n = C.c_int64()
        N.check(N.lib().sgz_ctx_trim(self._h, C.byref(n)))
        return n.value

    def measure_peak(self, which: int) -> float:
        v = C.c_double()
        N.check(N.lib().sgz_measure_peak(self._h, int(which), C.byref(v)))
        return v.value


def device_count() -> int:
    n = C.c_int32()
    N.check(N.lib().sgz_device_count(C.byref(n)))
    return n.value


class Database:
    """sgz_db: the feature database resident in HBM."""

    def __init__(self, ctx: Context, num_ch: int, norm: Optional[np.ndarray] = None):
        self.ctx = ctx
        self.num_ch = int(num_ch)
        self._h = C.c_void_p()
        nrm = None
        if norm is not None:
            nrm = np.ascontiguousarray(norm, np.float32)
            if nrm.shape != (num_ch, 2):
                raise ValueError(f"norm must have shape ({num_ch}, 2), got {nrm.shape}")  # reference: require
        N.check(N.lib().sgz_db_create(ctx._h, self.num_ch, None if nrm is None else N.fptr(nrm), C.byref(self._h)))

    def close(self):
        if self._h:
            N.lib().sgz_db_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, total_frames: int, num_files: int = 0):
        N.check(N.lib().sgz_db_reserve(self._h, C.c_int64(total_frames), int(num_files)))

    def add_file(self, frames: np.ndarray, layout: int = N.LAYOUT_INTERLEAVED_LE) -> int:
        if layout == N.LAYOUT_PLANAR_LE:
            a = np.ascontiguousarray(frames, np.float32)
            n = a.shape[1]
        else:
            a = _frames(frames) if layout == N.LAYOUT_INTERLEAVED_LE else np.ascontiguousarray(frames)
            n = a.shape[0]
        return N.check(N.lib().sgz_db_add_file(self._h, N.fptr(a), C.c_int64(n), int(layout)))

    def add_file_ptr(self, ptr: int, n_frames: int, layout: int = N.LAYOUT_INTERLEAVED_LE) -> int:
        return N.check(N.lib().sgz_db_add_file(self._h, C.c_void_p(ptr), C.c_int64(n_frames), int(layout)))

    def add_file_device(self, dptr: int, n_frames: int) -> int:
        return N.check(N.lib().sgz_db_add_file_device(self._h, C.c_void_p(dptr), C.c_int64(n_frames)))

    def add_synth(self, seed: int, stream: int, n_frames: int, mu: np.ndarray, sigma: np.ndarray,
                  floor0: float) -> int:
        mu = np.ascontiguousarray(mu, np.float32)
        sigma = np.ascontiguousarray(sigma, np.float32)
        return N.check(N.lib().sgz_db_add_synth(self._h, C.c_uint64(seed), C.c_uint32(stream), C.c_int64(n_frames),
                                                N.fptr(mu), N.fptr(sigma), C.c_float(floor0)))

    def add_synth_many(self, seed: int, first_stream: int, num_files: int, n_frames_each: int, mu: np.ndarray,
                       sigma: np.ndarray, floor0: float) -> int:
        """`num_files` synthetic files in one kernel launch (streams first_stream, first_stream + 1, ...)."""
        mu = np.ascontiguousarray(mu, np.float32)
        sigma = np.ascontiguousarray(sigma, np.float32)
        return N.check(N.lib().sgz_db_add_synth_many(self._h, C.c_uint64(seed), C.c_uint32(first_stream), int(num_files),
                                                     C.c_int64(n_frames_each), N.fptr(mu), N.fptr(sigma),
                                                     C.c_float(floor0)))

    def patch(self, file: int, frame_off: int, frames: np.ndarray):
        a = _frames(frames)
        N.check(N.lib().sgz_db_patch(self._h, int(file), C.c_int64(frame_off), N.fptr(a), C.c_int64(a.shape[0])))

    def finalize(self, wait: bool = True):
        """wait=False: do not wait for HOST_STABLE uploads in flight; the first search streams behind them."""
        N.check((N.lib().sgz_db_finalize if wait else N.lib().sgz_db_finalize_async)(self._h))

    def stats(self, want_per_file: bool = False):
        """FeatureStats over a RAW database (norm=None): [numCh][2] float64 (p01 min, p99 max)."""
        nf = self.info()[0]
        out = np.zeros((self.num_ch, 2), np.float64)
        per = np.zeros((max(nf, 1), self.num_ch, 2), np.float64) if want_per_file else None
        N.check(N.lib().sgz_db_stats(self._h, out.ctypes.data_as(C.c_void_p),
                                     None if per is None else per.ctypes.data_as(C.c_void_p)))
        return (out, per[:nf]) if want_per_file else out

    def info(self) -> Tuple[int, int, int]:
        nf, tf, nc = C.c_int32(), C.c_int64(), C.c_int32()
        N.check(N.lib().sgz_db_info(self._h, C.byref(nf), C.byref(tf), C.byref(nc)))
        return nf.value, tf.value, nc.value

    def file_frames(self, file: int) -> int:
        n = C.c_int64()
        N.check(N.lib().sgz_db_file_frames(self._h, int(file), C.byref(n)))
        return n.value

    def read(self, file: int, frame_off: int, n: int) -> np.ndarray:
        """Normalised frames as planar [numCh][n]."""
        out = np.empty((self.num_ch, n), np.float32)
        N.check(N.lib().sgz_db_read(self._h, int(file), C.c_int64(frame_off), C.c_int64(n), N.fptr(out)))
        return out


def _matches(arr, n) -> List[dict]:
    return [dict(sim=arr[i].sim, file=arr[i].file, start=arr[i].start, stop=arr[i].stop,
                 boostIn=arr[i].boostIn, boostOut=arr[i].boostOut) for i in range(n)]


class CorrelationJob:
    """sgz_corr: one FeatureCorrelation search over a Database (or over this rank's shard)."""

    def __init__(self, db: Database, cfg: N.CorrConfig, input_frames: np.ndarray,
                 layout: int = N.LAYOUT_INTERLEAVED_LE):
        self.db = db
        a = _frames(input_frames)
        self._h = C.c_void_p()
        N.check(N.lib().sgz_corr_create(db._h, C.byref(cfg), N.fptr(a), C.c_int64(a.shape[0]), int(layout),
                                        C.byref(self._h)))
        self.cfg = cfg
        # sharded searches: punch-in-only jobs exchange the numMatches largest file maxima per rank instead of one
        # summary per file (sgz_corr_local_top / sgz_corr_set_global_top)
        self.sparse_summary = not bool(cfg.hasPunchOut)
        # ... and with one match per file the whole search is ONE exchange (sgz_corr_local_best / finish_from_best)
        self.one_exchange = self.sparse_summary and int(cfg.numPerFile) == 1 and os.environ.get("SGZ_ONE_EXCHANGE", "1") != "0"

    def close(self):
        if self._h:
            N.lib().sgz_corr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- one-call ----
    def run(self) -> List[dict]:
        N.check(N.lib().sgz_corr_run(self._h))
        return self.result()

    def start(self):
        N.check(N.lib().sgz_corr_start(self._h))

    def poll(self) -> Tuple[float, bool, int]:
        p, d, s = C.c_float(), C.c_int32(), C.c_int32()
        N.check(N.lib().sgz_corr_poll(self._h, C.byref(p), C.byref(d), C.byref(s)))
        return p.value, bool(d.value), s.value

    def abort(self):
        N.check(N.lib().sgz_corr_abort(self._h))

    def wait(self):
        N.check(N.lib().sgz_corr_wait(self._h))

    def result(self) -> List[dict]:
        n = C.c_int32()
        N.check(N.lib().sgz_corr_result(self._h, None, 0, C.byref(n)))
        arr = (N.Match * max(n.value, 1))()
        N.check(N.lib().sgz_corr_result(self._h, arr, max(n.value, 1), C.byref(n)))
        return _matches(arr, n.value)

    @property
    def num_offsets(self) -> int:
        n = C.c_int64()
        N.check(N.lib().sgz_corr_num_offsets(self._h, C.byref(n)))
        return n.value

    def timing(self) -> dict:
        a, b, c = C.c_float(), C.c_float(), C.c_int64()
        N.check(N.lib().sgz_corr_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(scan_ms=a.value, select_ms=b.value, scan_launches=c.value)

    def curve(self, file: int, which: int = 0, first: int = 0, n: Optional[int] = None):
        if n is None:
            raise ValueError("n required")
        sim = np.empty(n, np.float32)
        boost = np.empty(n, np.float32)
        N.check(N.lib().sgz_corr_curve(self._h, int(which), int(file), C.c_int64(first), C.c_int64(n), N.fptr(sim),
                                       N.fptr(boost)))
        return sim, boost

    # ---- phase-wise protocol (sharded DB) ----
    def scan(self):
        N.check(N.lib().sgz_corr_scan(self._h))

    def local_summary(self) -> np.ndarray:
        n = C.c_int32()
        N.check(N.lib().sgz_corr_local_summary(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), N.SUMMARY_DTYPE)
        N.check(N.lib().sgz_corr_local_summary(self._h, N.fptr(out), out.shape[0], C.byref(n)))
        return out[:n.value]

    def set_global(self, all_summaries: np.ndarray, my_first_file: int):
        a = np.ascontiguousarray(all_summaries, N.SUMMARY_DTYPE)
        N.check(N.lib().sgz_corr_set_global(self._h, N.fptr(a), a.shape[0], int(my_first_file)))

    def local_top(self):
        """(entries with LOCAL file indices, number of local files): the sparse summary of a punch-in-only search"""
        n, nf = C.c_int32(), C.c_int32()
        N.check(N.lib().sgz_corr_local_top(self._h, None, 0, C.byref(n), C.byref(nf)))
        out = np.zeros(max(n.value, 1), N.ENTRY_DTYPE)
        N.check(N.lib().sgz_corr_local_top(self._h, N.fptr(out), out.shape[0], C.byref(n), C.byref(nf)))
        return out[:n.value], nf.value

    def set_global_top(self, all_entries: np.ndarray, n_files_global: int, my_first_file: int):
        a = np.ascontiguousarray(all_entries, N.ENTRY_DTYPE)
        N.check(N.lib().sgz_corr_set_global_top(self._h, N.fptr(a) if a.shape[0] else None, a.shape[0], int(n_files_global),
                                                int(my_first_file)))

    def local_best(self):
        """(records of this rank's numMatches best files with LOCAL file indices, number of local files, ok): the ONE
        message of a punch-in search with numPerFile = 1 (sgz_corr_local_best); ok = False -> use local_top / select"""
        n, nf, ok = C.c_int32(), C.c_int32(), C.c_int32()
        N.check(N.lib().sgz_corr_local_best(self._h, None, 0, C.byref(n), C.byref(nf), C.byref(ok)))
        out = np.zeros(max(n.value, 1), N.RECORD_DTYPE)
        N.check(N.lib().sgz_corr_local_best(self._h, N.fptr(out), out.shape[0], C.byref(n), C.byref(nf), C.byref(ok)))
        return out[:n.value], nf.value, bool(ok.value)

    def finish_from_best(self, all_records: np.ndarray, n_files_global: int):
        a = np.ascontiguousarray(all_records, N.RECORD_DTYPE)
        N.check(N.lib().sgz_corr_finish_from_best(self._h, N.fptr(a) if a.shape[0] else None, a.shape[0], int(n_files_global)))

    def select(self) -> np.ndarray:
        n = C.c_int32()
        N.check(N.lib().sgz_corr_select(self._h, C.byref(n)))
        out = np.zeros(max(n.value, 1), N.RECORD_DTYPE)
        N.check(N.lib().sgz_corr_records(self._h, N.fptr(out), out.shape[0], C.byref(n)))
        return out[:n.value]

    def merge(self, all_records: np.ndarray) -> bool:
        a = np.ascontiguousarray(all_records, N.RECORD_DTYPE)
        done = C.c_int32()
        N.check(N.lib().sgz_corr_merge(self._h, N.fptr(a) if a.shape[0] else None, a.shape[0], C.byref(done)))
        return bool(done.value)


def segm_run(ctx: Context, cfg: N.SegmConfig, frames: np.ndarray, norm: Optional[np.ndarray] = None,
             want_curve: bool = False):
    a = _frames(frames)
    nrm = None if norm is None else np.ascontiguousarray(norm, np.float32)
    cap = max(cfg.numBreaks, 0) + 1
    out = (N.Break * cap)()
    n, noff = C.c_int32(), C.c_int64()
    curve = np.full(max(a.shape[0], 1), np.nan, np.float32) if want_curve else None
    N.check(N.lib().sgz_segm_run(ctx._h, C.byref(cfg), a.shape[1], None if nrm is None else N.fptr(nrm), N.fptr(a),
                                 C.c_int64(a.shape[0]), N.LAYOUT_INTERLEAVED_LE, out, cap, C.byref(n),
                                 None if curve is None else N.fptr(curve),
                                 C.c_int64(0 if curve is None else curve.shape[0]), C.byref(noff)))
    breaks = [dict(sim=out[i].sim, pos=out[i].pos) for i in range(n.value)]
    if want_curve:
        return breaks, curve[:noff.value], noff.value
    return breaks


def cross_run(ctx: Context, cfg: N.CrossConfig, frames1: np.ndarray, frames2: np.ndarray,
              norm: Optional[np.ndarray] = None) -> np.ndarray:
    """CrossSimilarity: the sim curve the reference writes into its 1-channel output file."""
    a1, a2 = _frames(frames1), _frames(frames2)
    nrm = None if norm is None else np.ascontiguousarray(norm, np.float32)
    n = C.c_int64()
    N.check(N.lib().sgz_cross_num_outputs(C.byref(cfg), C.c_int64(a1.shape[0]), C.c_int64(a2.shape[0]), C.byref(n)))
    sim = np.zeros(max(n.value, 1), np.float32)
    N.check(N.lib().sgz_cross_run(ctx._h, C.byref(cfg), a1.shape[1], None if nrm is None else N.fptr(nrm),
                                  N.fptr(a1), C.c_int64(a1.shape[0]), N.fptr(a2), C.c_int64(a2.shape[0]),
                                  N.LAYOUT_INTERLEAVED_LE, N.fptr(sim), C.c_int64(sim.shape[0]), C.byref(n)))
    return sim[:n.value]


def self_geometry(cfg: N.SelfConfig, n1: int, n2: int) -> dict:
    g = N.SelfGeometry()
    N.check(N.lib().sgz_self_geometry_of(C.byref(cfg), C.c_int64(n1), C.c_int64(n2), C.byref(g)))
    return dict(imgExt=g.imgExt, decim=g.decim, numCorrs=g.numCorrs, afStart=g.afStart, numCells=g.numCells)


SELF_KERNELS = {0: "none", 1: "ffma2_gram", 2: "tc_gram", 3: "fp64_replay"}


def self_last_kernel(ctx: Context) -> str:
    """Which kernel rendered the last SelfSimilarity image of this context (sgz_self_last_kernel)."""
    return SELF_KERNELS[int(N.lib().sgz_self_last_kernel(ctx._h))]


def self_run(ctx: Context, cfg: N.SelfConfig, frames1: np.ndarray, frames2: Optional[np.ndarray] = None,
             norm: Optional[np.ndarray] = None, col_begin: int = 0, col_end: int = 0, download: bool = True):
    a1 = _frames(frames1)
    a2 = None if frames2 is None else _frames(frames2)
    nrm = None if norm is None else np.ascontiguousarray(norm, np.float32)
    g = self_geometry(cfg, a1.shape[0], a1.shape[0] if a2 is None else a2.shape[0])
    ext = g["imgExt"]
    rgb = np.zeros((max(ext, 1), max(ext, 1)), np.int32) if download else None
    geom = N.SelfGeometry()
    N.check(N.lib().sgz_self_run(ctx._h, C.byref(cfg), a1.shape[1], None if nrm is None else N.fptr(nrm), N.fptr(a1),
                                 C.c_int64(a1.shape[0]), None if a2 is None else N.fptr(a2),
                                 C.c_int64(0 if a2 is None else a2.shape[0]), N.LAYOUT_INTERLEAVED_LE,
                                 int(col_begin), int(col_end), None if rgb is None else N.fptr(rgb),
                                 C.c_int64(0 if rgb is None else rgb.size), C.byref(geom)))
    return (rgb[:ext, :ext] if rgb is not None else None), g


def self_cells(ctx: Context, cfg: N.SelfConfig, frames1: np.ndarray, frames2: Optional[np.ndarray],
               left: np.ndarray, right: np.ndarray, norm: Optional[np.ndarray] = None):
    a1 = _frames(frames1)
    a2 = None if frames2 is None else _frames(frames2)
    nrm = None if norm is None else np.ascontiguousarray(norm, np.float32)
    l = np.ascontiguousarray(left, np.int32)
    r = np.ascontiguousarray(right, np.int32)
    sim = np.empty(l.shape[0], np.float32)
    rgb = np.empty(l.shape[0], np.int32)
    N.check(N.lib().sgz_self_cells(ctx._h, C.byref(cfg), a1.shape[1], None if nrm is None else N.fptr(nrm),
                                   N.fptr(a1), C.c_int64(a1.shape[0]), None if a2 is None else N.fptr(a2),
                                   C.c_int64(0 if a2 is None else a2.shape[0]), N.LAYOUT_INTERLEAVED_LE,
                                   C.c_int64(l.shape[0]), N.fptr(l), N.fptr(r), N.fptr(sim), N.fptr(rgb)))
    return sim, rgb
