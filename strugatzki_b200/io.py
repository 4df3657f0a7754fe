"""Host-side file formats that cross the drop-in boundary:

* `Span` -- the subset of de.sciss.span the hot path uses (Span(start, stop), Span.from, Span.until, Span.all).
* feature files: float32 big-endian AIFF-C ('fl32'), numCoeffs+1 channels, channel 0 = loudness
  (written by ScalaAudioFile: NonRealtimeProcessor.scala:164-165); `feat_norms.aif` = 2 frames {min, max}
  (Strugatzki.scala:417-426).
* `<feature>` meta XML of FeatureExtraction.Config (Api/FeatureExtraction.scala:163-206).
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, replace
from typing import List, Optional, Tuple

import numpy as np

NORMALIZE_NAME = "feat_norms.aif"   # Strugatzki.NormalizeName, Strugatzki.scala:36


# ---------------------------------------------------------------------------------------------------------
# Span
# ---------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Span:
    """de.sciss.span.Span.NonVoid: start / stop in sample frames, either may be open."""
    start: Optional[int] = None
    stop: Optional[int] = None

    @staticmethod
    def all() -> "Span":
        return Span(None, None)

    @staticmethod
    def from_(start: int) -> "Span":
        return Span(int(start), None)

    @staticmethod
    def until(stop: int) -> "Span":
        return Span(None, int(stop))

    @property
    def has_start(self) -> bool:
        return self.start is not None

    @property
    def has_stop(self) -> bool:
        return self.stop is not None

    @property
    def length(self) -> int:
        return self.stop - self.start

    def to_xml(self) -> List[ET.Element]:
        out = []
        if self.has_start:
            e = ET.Element("start"); e.text = str(self.start); out.append(e)
        if self.has_stop:
            e = ET.Element("stop"); e.text = str(self.stop); out.append(e)
        return out

    @staticmethod
    def from_xml(node: Optional[ET.Element]) -> "Span":   # SpanUtil.fromXML, SpanUtil.scala:21-30
        if node is None:
            return Span.all()
        s, t = node.find("start"), node.find("stop")
        return Span(None if s is None else int(s.text), None if t is None else int(t.text))


def span_spacing(a: Span, b: Span) -> int:   # SpanUtil.spacing, SpanUtil.scala:38-43
    return b.start - a.stop if a.start < b.start else a.start - b.stop


# ---------------------------------------------------------------------------------------------------------
# AIFF / AIFF-C float32
# ---------------------------------------------------------------------------------------------------------
def _ext80(x: float) -> bytes:
    """IEEE 754 80-bit extended (AIFF sample rate)."""
    if x == 0:
        return b"\0" * 10
    import math
    sign = 0x8000 if x < 0 else 0
    m, e = math.frexp(abs(x))
    e += 16382
    mant = int(m * (1 << 64))
    return struct.pack(">HQ", sign | e, mant)


def _from_ext80(b: bytes) -> float:
    e, mant = struct.unpack(">HQ", b)
    sign = -1.0 if e & 0x8000 else 1.0
    e &= 0x7FFF
    if e == 0 and mant == 0:
        return 0.0
    return sign * mant * 2.0 ** (e - 16383 - 63)


def write_aiff(path: str, frames: np.ndarray, sample_rate: float = 44100.0) -> None:
    """float32 AIFF-C ('fl32'); frames = [numFrames][numChannels]."""
    a = np.ascontiguousarray(frames, np.float32)
    n, ch = a.shape
    data = a.astype(">f4").tobytes()
    comp = b"fl32" + bytes([0x0C]) + b"Float 32-bit" + b"\0"        # pstring, padded to even length
    comm = struct.pack(">hIh", ch, n, 32) + _ext80(sample_rate) + comp
    fver = struct.pack(">I", 0xA2805140)
    ssnd = struct.pack(">II", 0, 0) + data
    chunks = b"FVER" + struct.pack(">I", 4) + fver + b"COMM" + struct.pack(">I", len(comm)) + comm + \
             b"SSND" + struct.pack(">I", len(ssnd)) + ssnd + (b"\0" if len(ssnd) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"FORM" + struct.pack(">I", 4 + len(chunks)) + b"AIFC" + chunks)


@dataclass
class AudioFileSpec:
    num_channels: int
    num_frames: int
    sample_rate: float
    big_endian_f32: bool   # payload can be handed to the engine as SGZ_LAYOUT_INTERLEAVED_BE


def read_aiff(path: str, raw: bool = False):
    """Reads an AIFF / AIFF-C file.  Returns (frames [n][ch] float32, spec); with raw=True the float32
    payload is returned still big-endian (dtype '>f4') so the GPU does the byte swap."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"FORM" or buf[8:12] not in (b"AIFF", b"AIFC"):
        raise IOError(f"{path}: not an AIFF file")
    pos, comm, ssnd = 12, None, None
    while pos + 8 <= len(buf):
        cid, size = buf[pos:pos + 4], struct.unpack(">I", buf[pos + 4:pos + 8])[0]
        body = buf[pos + 8:pos + 8 + size]
        if cid == b"COMM":
            comm = body
        elif cid == b"SSND":
            ssnd = body
        pos += 8 + size + (size & 1)
    if comm is None:
        raise IOError(f"{path}: missing COMM chunk")
    ch, n, bits = struct.unpack(">hIh", comm[:8])
    sr = _from_ext80(comm[8:18])
    ctype = comm[18:22] if len(comm) >= 22 else b"NONE"
    if n == 0:
        return np.zeros((0, ch), np.float32), AudioFileSpec(ch, 0, sr, True)
    if ssnd is None:
        raise IOError(f"{path}: missing SSND chunk")
    off = struct.unpack(">I", ssnd[:4])[0]
    payload = ssnd[8 + off:]
    if ctype in (b"fl32", b"FL32"):
        a = np.frombuffer(payload, ">f4", n * ch).reshape(n, ch)
        return (a if raw else a.astype(np.float32)), AudioFileSpec(ch, n, sr, True)
    if ctype in (b"fl64", b"FL64"):
        a = np.frombuffer(payload, ">f8", n * ch).reshape(n, ch).astype(np.float32)
        return a, AudioFileSpec(ch, n, sr, False)
    if ctype == b"NONE":
        if bits == 16:
            a = np.frombuffer(payload, ">i2", n * ch).astype(np.float32) / 32768.0
        elif bits == 32:
            a = np.frombuffer(payload, ">i4", n * ch).astype(np.float32) / 2147483648.0
        elif bits == 24:
            b3 = np.frombuffer(payload, np.uint8, n * ch * 3).reshape(-1, 3).astype(np.int32)
            v = (b3[:, 0] << 16) | (b3[:, 1] << 8) | b3[:, 2]
            v = np.where(v & 0x800000, v - 0x1000000, v)
            a = v.astype(np.float32) / 8388608.0
        else:
            raise IOError(f"{path}: unsupported PCM width {bits}")
        return a.reshape(n, ch).astype(np.float32), AudioFileSpec(ch, n, sr, False)
    raise IOError(f"{path}: unsupported AIFF-C compression {ctype!r}")


def read_aiff_many(paths, raw: bool = True, workers: int = 8, window: int = 16):
    """Ingest pipeline for a database folder: the feature files are read and parsed by a pool
    of `workers` threads up to `window` files ahead of the consumer (file I/O and numpy's frombuffer release the GIL) and
    yielded IN THE ORDER of `paths` as (frames, spec) -- the order is what fixes the file indices of the search
    (FeatureCorrelationImpl.scala:160-165).  The consumer (sgz_db_add_file, which copies into a pinned staging ring and
    returns while the H2D copy is in flight) therefore overlaps disk, parse and upload.  A failed read raises at the
    position of its file; closing the generator early (abort) cancels the reads not yet started."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor

    paths = list(paths)
    if workers <= 1 or len(paths) <= 1:
        for path in paths:
            yield read_aiff(path, raw=raw)
        return
    window = max(window, workers)
    pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="sgz-ingest")
    pending = deque()
    try:
        it = iter(paths)
        for path in it:
            pending.append(pool.submit(read_aiff, path, raw))
            if len(pending) >= window:
                break
        while pending:
            fut = pending.popleft()
            res = fut.result()
            nxt = next(it, None)
            if nxt is not None:
                pending.append(pool.submit(read_aiff, nxt, raw))
            yield res
    finally:
        for fut in pending:
            fut.cancel()
        pool.shutdown(wait=True, cancel_futures=True)


def read_norm_file(database_folder: str, num_ch: int) -> np.ndarray:
    """feat_norms.aif -> [numCh][2] = {min, max}; same `require` as FeatureCorrelationImpl.scala:61-71."""
    a, spec = read_aiff(os.path.join(database_folder, NORMALIZE_NAME))
    if not (spec.num_channels == num_ch and spec.num_frames == 2):
        raise ValueError(f"requirement failed: {NORMALIZE_NAME} must have {num_ch} channels x 2 frames, "
                         f"has {spec.num_channels} x {spec.num_frames}")
    return np.ascontiguousarray(a.T, np.float32)   # AudioFile.buffer layout: [channel][frame]


def write_norm_file(database_folder: str, norm: np.ndarray) -> None:
    write_aiff(os.path.join(database_folder, NORMALIZE_NAME), np.ascontiguousarray(norm, np.float32).T)


# ---------------------------------------------------------------------------------------------------------
# FeatureExtraction.Config meta XML
# ---------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FeatureExtractionConfig:
    """Api/FeatureExtraction.scala:58-172 (only the persisted fields; the extractor itself is out of scope)."""
    audio_input: str = "input.aif"
    feature_output: str = "features.aif"
    meta_output: Optional[str] = None
    num_coeffs: int = 13
    fft_size: int = 1024
    fft_overlap: int = 2
    channels_behavior: int = 0    # Mix = 0, First = 1, Last = 2

    @property
    def step_size(self) -> int:
        return self.fft_size // self.fft_overlap

    def to_xml(self) -> ET.Element:
        root = ET.Element("feature")
        for tag, val in (("input", self.audio_input), ("output", self.feature_output),
                         ("meta", self.meta_output or ""), ("numCoeffs", self.num_coeffs),
                         ("fftSize", self.fft_size), ("fftOverlap", self.fft_overlap),
                         ("channels", self.channels_behavior)):
            e = ET.SubElement(root, tag)
            e.text = str(val)
        return root

    @staticmethod
    def from_xml(root: ET.Element) -> "FeatureExtractionConfig":
        def txt(tag):
            e = root.find(tag)
            return "" if e is None or e.text is None else e.text
        ch = txt("channels")
        if ch and int(ch) not in (0, 1, 2):
            raise ValueError(ch)           # ChannelsBehavior.apply -> IllegalArgumentException
        return FeatureExtractionConfig(txt("input"), txt("output"), txt("meta") or None, int(txt("numCoeffs")),
                                       int(txt("fftSize")), int(txt("fftOverlap")), int(ch) if ch else 0)

    @staticmethod
    def from_xml_file(path: str) -> "FeatureExtractionConfig":
        try:
            return FeatureExtractionConfig.from_xml(ET.parse(path).getroot())
        except ET.ParseError as e:
            raise IOError(f"In file: {path}") from e

    def write(self, path: str) -> None:
        ET.ElementTree(self.to_xml()).write(path, encoding="utf-8", xml_declaration=True)


def full_to_feat(n: int, step: int) -> int:
    return int((n + (step >> 1)) // step)


def list_database(database_folder: str, meta_input: str, num_coeffs: int, step_size: int) -> List[FeatureExtractionConfig]:
    """DB discovery of FeatureCorrelationImpl.scala:42-55.  The reference iterates a HashSet of java.io.File
    (arbitrary but fixed order); here the order is the sorted file name, which is what `Match.file` indices
    refer to."""
    out = []
    meta_abs = os.path.abspath(meta_input)
    for name in sorted(os.listdir(database_folder)):
        if not name.endswith("_feat.xml"):
            continue
        p = os.path.join(database_folder, name)
        if os.path.abspath(p) == meta_abs:
            continue
        e = FeatureExtractionConfig.from_xml_file(p)
        if e.num_coeffs == num_coeffs and e.fft_size // e.fft_overlap == step_size:
            out.append(e)
    return out


class DatabaseCache:
    """On-disk cache of a decoded feature database: `<folder>/.sgz_cache/db_<key>.f32` holds
    the RAW float32 little-endian interleaved frames of all database files back to back, `db_<key>.json` their frame
    counts.  The key covers name, size and mtime of every feature file and the channel count, so any change of the
    database makes a new cache; normalisation stays on the GPU, so the norm file is not part of the key.  A search then
    maps one file instead of parsing thousands of AIFFs (FeatureCorrelationImpl.scala:169,195 is what that replaces), and
    the mapped frames are contiguous, so the engine uploads them in 32 MB copies.  SGZ_DB_CACHE=0 turns it off."""

    def __init__(self, folder: str, feature_files: List[str], num_ch: int):
        import hashlib
        self.num_ch = num_ch
        self.files = list(feature_files)
        h = hashlib.sha1(f"sgz-db-cache-1 {num_ch}".encode())
        for f in self.files:
            st = os.stat(f)
            h.update(f"{os.path.basename(f)} {st.st_size} {st.st_mtime_ns}\n".encode())
        self.dir = os.path.join(folder, ".sgz_cache")
        self.base = os.path.join(self.dir, f"db_{h.hexdigest()[:20]}")
        self.enabled = os.environ.get("SGZ_DB_CACHE", "1") != "0"

    def load(self):
        """(memmap of float32 [total frames][num_ch], frame counts) or None"""
        if not self.enabled:
            return None
        try:
            import json
            with open(self.base + ".json") as fh:
                meta = json.load(fh)
            counts = [int(c) for c in meta["frames"]]
            if len(counts) != len(self.files) or meta["num_ch"] != self.num_ch:
                return None
            total = sum(counts)
            if os.path.getsize(self.base + ".f32") != total * self.num_ch * 4:
                return None
            mm = np.memmap(self.base + ".f32", np.float32, "r", shape=(max(total, 1), self.num_ch)) if total else \
                np.zeros((0, self.num_ch), np.float32)
            return mm, counts
        except (OSError, ValueError, KeyError):
            return None

    def writer(self):
        """context manager with add(frames) -- frames as decoded by read_aiff_many(raw=True): big-endian payloads are
        swapped here, once -- that publishes the cache atomically on exit; any I/O error just leaves no cache"""
        cache = self

        class _W:
            def __enter__(self):
                self.counts, self.fh = [], None
                if cache.enabled:
                    try:
                        os.makedirs(cache.dir, exist_ok=True)
                        self.fh = open(cache.base + ".f32.tmp", "wb")
                    except OSError:
                        self.fh = None
                return self

            def add(self, frames: np.ndarray):
                self.counts.append(int(frames.shape[0]))
                if self.fh is not None:
                    try:
                        self.fh.write(np.ascontiguousarray(frames, "<f4").tobytes())
                    except OSError:
                        self.fh.close()
                        self.fh = None

            def __exit__(self, et, ev, tb):
                if self.fh is None:
                    return False
                self.fh.close()
                try:
                    if et is None and len(self.counts) == len(cache.files):
                        import json
                        with open(cache.base + ".json.tmp", "w") as fh:
                            json.dump({"num_ch": cache.num_ch, "frames": self.counts,
                                       "files": [os.path.basename(f) for f in cache.files]}, fh)
                        os.replace(cache.base + ".f32.tmp", cache.base + ".f32")
                        os.replace(cache.base + ".json.tmp", cache.base + ".json")
                    else:
                        os.remove(cache.base + ".f32.tmp")
                except OSError:
                    pass
                return False

        return _W()
