"""FeatureStats -- host-side mirror of FeatureStats.scala with a body on the B200 engine (replaces
Impl/FeatureStatsImpl.scala:30-135).  Config = the list of feature AIFF paths,
Product = per channel (min over files of the 1st percentile, max over files of the 99th) as Doubles."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np

from . import _native as N
from . import engine
from .io import read_aiff_many, write_norm_file
from .processor import Aborted, ProcessorFactory, ProcessorImpl


class FeatureStatsImpl(ProcessorImpl):
    device = 0

    def body(self) -> List[Tuple[float, float]]:
        paths: Sequence[str] = list(self.config)
        if not paths:
            raise ValueError("requirement failed: no feature files")      # allMins stays null upstream -> NPE
        ctx = engine.Context(self.device)
        try:
            db = None
            for i, (frames, spec) in enumerate(read_aiff_many(paths, raw=True)):   # reader threads run ahead of the upload
                self.check_aborted()                              # (big-endian payload: the GPU swaps the bytes)
                if db is None:
                    db = engine.Database(ctx, spec.num_channels, None)
                elif spec.num_channels != db.num_ch:
                    raise ValueError("requirement failed")        # :43
                if frames.dtype == np.dtype(">f4"):
                    db.add_file(frames, N.LAYOUT_INTERLEAVED_BE)
                else:
                    db.add_file(frames)
                self.progress = 0.5 * (i + 1) / len(paths)
            db.finalize()
            out = db.stats()
            db.close()
        except N.Aborted:
            raise Aborted()
        finally:
            ctx.close()
        self.progress = 1.0
        return [(float(lo), float(hi)) for lo, hi in out]


class FeatureStats(ProcessorFactory):
    Impl = FeatureStatsImpl

    @classmethod
    def default_config(cls):
        return []


def write_norms(database_folder: str, spans: Sequence[Tuple[float, float]]) -> None:
    """what `Strugatzki --stats` does with the product (Strugatzki.scala:414-426): feat_norms.aif, one channel per
    feature, frame 0 = min, frame 1 = max, narrowed to Float"""
    write_norm_file(database_folder, np.asarray(spans, np.float64).astype(np.float32))


def stats_of_folder(database_folder: str, observer=None) -> List[Tuple[float, float]]:
    """`Strugatzki --stats -d <dir>`: all *_feat.aif of the folder (Strugatzki.scala:411-413)"""
    paths = sorted(os.path.join(database_folder, n) for n in os.listdir(database_folder) if n.endswith("_feat.aif"))
    res = FeatureStats.run(paths, observer).await_result()
    write_norms(database_folder, res)
    return res
