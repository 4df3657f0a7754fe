"""FeatureCorrelation -- host-side mirror of Api/FeatureCorrelation.scala (Config / Punch / Match / XML) with
a processor body that runs on the B200 engine instead of Impl/FeatureCorrelationImpl.scala.

Field names, defaults and XML tags are the reference's (Api/FeatureCorrelation.scala:105-273); all durations
are sample frames of the original audio.  `Match.file` is the DB entry's audio input path, as in the
reference (:237).
"""
from __future__ import annotations

import time
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field, replace
from typing import List, Optional

import numpy as np

from . import _native as N
from . import engine
from .io import (DatabaseCache, FeatureExtractionConfig, Span, list_database, read_aiff, read_aiff_many, read_norm_file)
from .processor import Aborted, ProcessorFactory, ProcessorImpl

verbose = False


@dataclass(frozen=True)
class Punch:
    """Api/FeatureCorrelation.scala:93-100"""
    span: Span
    temporalWeight: float = 0.5

    def to_xml(self, tag: str = "punch") -> ET.Element:
        e = ET.Element(tag)
        for c in self.span.to_xml():
            e.append(c)
        w = ET.SubElement(e, "weight")
        w.text = repr(float(np.float32(self.temporalWeight)))
        return e

    @staticmethod
    def from_xml(e: ET.Element) -> "Punch":
        return Punch(Span(int(e.find("start").text), int(e.find("stop").text)),
                     float(np.float32(float(e.find("weight").text))))


@dataclass(frozen=True)
class Match:
    """Api/FeatureCorrelation.scala:54-70"""
    sim: float
    file: str
    punch: Span
    boostIn: float
    boostOut: float

    def to_xml(self) -> ET.Element:
        e = ET.Element("match")
        for tag, v in (("sim", repr(float(self.sim))), ("file", self.file), ("start", self.punch.start),
                       ("stop", self.punch.stop), ("boostIn", repr(float(self.boostIn))),
                       ("boostOut", repr(float(self.boostOut)))):
            c = ET.SubElement(e, tag)
            c.text = str(v)
        return e

    @staticmethod
    def from_xml(e: ET.Element) -> "Match":
        return Match(float(e.find("sim").text), e.find("file").text, Span(int(e.find("start").text),
                     int(e.find("stop").text)), float(e.find("boostIn").text), float(e.find("boostOut").text))

    def pretty(self) -> str:
        return (f"Match(\n   sim      = {self.sim}\n   file     = {self.file}\n   punch    = {self.punch}"
                f"\n   boostIn  = {self.boostIn}\n   boostOut = {self.boostOut}\n)")


@dataclass(frozen=True)
class Config:
    """Immutable FeatureCorrelation.Config (case class Impl, :225-245)."""
    databaseFolder: str = "database"
    metaInput: str = "input_feat.xml"
    punchIn: Punch = Punch(Span(0, 44100), 0.5)
    punchOut: Optional[Punch] = None
    minPunch: int = 22050
    maxPunch: int = 88200
    normalize: bool = True
    maxBoost: float = 8.0
    numMatches: int = 1
    numPerFile: int = 1
    minSpacing: int = 0

    def to_xml(self) -> ET.Element:
        r = ET.Element("correlate")
        ET.SubElement(r, "database").text = self.databaseFolder
        ET.SubElement(r, "input").text = self.metaInput
        r.append(self.punchIn.to_xml("punchIn"))
        if self.punchOut is not None:
            r.append(self.punchOut.to_xml("punchOut"))
        for tag, v in (("minPunch", self.minPunch), ("maxPunch", self.maxPunch),
                       ("normalize", str(self.normalize).lower()), ("maxBoost", repr(float(self.maxBoost))),
                       ("numMatches", self.numMatches), ("numPerFile", self.numPerFile),
                       ("minSpacing", self.minSpacing)):
            ET.SubElement(r, tag).text = str(v)
        return r

    @staticmethod
    def from_xml(r: ET.Element) -> "Config":
        po = r.find("punchOut")
        return Config(r.find("database").text, r.find("input").text, Punch.from_xml(r.find("punchIn")),
                      None if po is None else Punch.from_xml(po), int(r.find("minPunch").text),
                      int(r.find("maxPunch").text), r.find("normalize").text.strip().lower() == "true",
                      float(r.find("maxBoost").text), int(r.find("numMatches").text),
                      int(r.find("numPerFile").text), int(r.find("minSpacing").text))

    @staticmethod
    def from_xml_file(path: str) -> "Config":
        return Config.from_xml(ET.parse(path).getroot())

    def pretty(self) -> str:
        return ("Settings(\n" + "".join(f"   {k:<14} = {getattr(self, k)}\n" for k in (
            "databaseFolder", "metaInput", "punchIn", "punchOut", "minPunch", "maxPunch", "normalize", "maxBoost",
            "numMatches", "numPerFile", "minSpacing")) + ")")


class ConfigBuilder:
    """Mutable builder with the reference's defaults (:168-223); `build()` freezes it."""

    def __init__(self, config: Optional[Config] = None):
        self.read(config or Config())

    def read(self, c: Config):
        for k in c.__dataclass_fields__:
            setattr(self, k, getattr(c, k))

    def build(self) -> Config:
        return Config(**{k: getattr(self, k) for k in Config.__dataclass_fields__})


def native_config(c: Config, step_size: int) -> N.CorrConfig:
    po = c.punchOut
    return N.CorrConfig(step_size, c.punchIn.span.start, c.punchIn.span.stop, c.punchIn.temporalWeight,
                        0 if po is None else 1, 0 if po is None else po.span.start, 0 if po is None else po.span.stop,
                        0.5 if po is None else po.temporalWeight, c.minPunch, c.maxPunch, c.maxBoost, c.numMatches,
                        c.numPerFile, c.minSpacing)


class FeatureCorrelationImpl(ProcessorImpl):
    """body() of the search; `device` selects the GPU (the reference has no such notion)."""
    device = 0

    def __init__(self, config):
        super().__init__(config.build() if isinstance(config, ConfigBuilder) else config)
        self._job = None
        self._cache_keepalive = None

    def _on_abort(self):
        if self._job is not None:
            self._job.abort()

    def body(self) -> List[Match]:
        cfg: Config = self.config
        extr_in = FeatureExtractionConfig.from_xml_file(cfg.metaInput)          # :35
        step = extr_in.step_size                                                # :36
        extr_dbs = list_database(cfg.databaseFolder, cfg.metaInput, extr_in.num_coeffs, step)   # :42-55
        if verbose:
            print(f"Number of compatible files in database : {len(extr_dbs)}")
        num_ch = extr_in.num_coeffs + 1
        norm = read_norm_file(cfg.databaseFolder, num_ch) if cfg.normalize else None            # :61-71
        inp, _ = read_aiff(extr_in.feature_output)
        self.check_aborted()
        ctx = engine.Context(self.device)
        db = engine.Database(ctx, num_ch, norm)
        try:
            # file order = order of extr_dbs (it fixes the file indices).  A database seen before comes from the on-disk
            # cache of its decoded frames (one mapped file, contiguous -> uploaded in 32 MB copies); otherwise reader
            # threads parse the AIFFs ahead of the upload and the cache is written on the way.
            cache = DatabaseCache(cfg.databaseFolder, [e.feature_output for e in extr_dbs], num_ch)
            hit = cache.load()
            if hit is not None:
                mm, counts = hit
                self._cache_keepalive = mm
                off = 0
                for n in counts:
                    self.check_aborted()
                    db.add_file_ptr(mm.ctypes.data + off * num_ch * 4, n, N.LAYOUT_INTERLEAVED_LE | N.LAYOUT_HOST_STABLE)
                    off += n
            else:
                with cache.writer() as w:
                    for e, (frames, spec) in zip(extr_dbs, read_aiff_many([e.feature_output for e in extr_dbs], raw=True)):
                        self.check_aborted()
                        if spec.num_channels != num_ch:
                            raise IOError(f"{e.feature_output}: {spec.num_channels} channels, expected {num_ch}")
                        if spec.big_endian_f32:
                            db.add_file(frames, N.LAYOUT_INTERLEAVED_BE)   # byte swap happens on the GPU
                        else:
                            db.add_file(frames, N.LAYOUT_INTERLEAVED_LE)
                        w.add(frames)
            db.finalize()
            self._job = engine.CorrelationJob(db, native_config(cfg, step), inp)
            self._job.start()
            while True:
                p, done, status = self._job.poll()
                self.progress = min(p, 0.999)
                if done:
                    break
                if self.aborted:
                    self._job.abort()
                time.sleep(0.0005)
            if status == N.ERR_ABORTED or self.aborted:
                raise Aborted()
            self._job.wait()
            res = self._job.result()
        except N.Aborted:
            raise Aborted()
        finally:
            if self._job is not None:
                self._job.close()
            db.close()
            ctx.close()
        self.progress = 1.0
        return [Match(m["sim"], extr_dbs[m["file"]].audio_input, Span(m["start"], m["stop"]), m["boostIn"],
                      m["boostOut"]) for m in res]


class FeatureCorrelation(ProcessorFactory):
    """object FeatureCorrelation extends ProcessorFactory.WithDefaults (:27-81)"""
    Impl = FeatureCorrelationImpl
    Config = Config
    ConfigBuilder = ConfigBuilder
    Punch = Punch
    Match = Match

    @classmethod
    def default_config(cls):
        return Config()
