"""FeatureSegmentation -- host-side mirror of Api/FeatureSegmentation.scala with a body on the B200 engine
(replaces Impl/FeatureSegmentationImpl.scala:31-142).  XML tags / defaults as in :134-191."""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native as N
from . import engine
from .io import FeatureExtractionConfig, Span, read_aiff, read_norm_file
from .processor import Aborted, ProcessorFactory, ProcessorImpl

verbose = False


@dataclass(frozen=True)
class Break:
    """Api/FeatureSegmentation.scala:47-55"""
    sim: float
    pos: int

    def to_xml(self) -> ET.Element:
        e = ET.Element("break")
        ET.SubElement(e, "sim").text = repr(float(self.sim))
        ET.SubElement(e, "pos").text = str(self.pos)
        return e

    @staticmethod
    def from_xml(e: ET.Element) -> "Break":
        return Break(float(e.find("sim").text), int(e.find("pos").text))

    def pretty(self) -> str:
        return f"Break(sim = {self.sim}, pos = {self.pos})"


@dataclass(frozen=True)
class Config:
    databaseFolder: str = "database"
    metaInput: str = "input_feat.xml"
    span: Span = Span.all()
    corrLen: int = 22050
    temporalWeight: float = 0.5
    normalize: bool = True
    numBreaks: int = 1
    minSpacing: int = 22050

    def to_xml(self) -> ET.Element:
        r = ET.Element("segmentation")
        ET.SubElement(r, "database").text = self.databaseFolder
        ET.SubElement(r, "input").text = self.metaInput
        sp = ET.SubElement(r, "span")
        for c in self.span.to_xml():
            sp.append(c)
        for tag, v in (("corr", self.corrLen), ("weight", repr(float(np.float32(self.temporalWeight)))),
                       ("normalize", str(self.normalize).lower()), ("numBreaks", self.numBreaks),
                       ("minSpacing", self.minSpacing)):
            ET.SubElement(r, tag).text = str(v)
        return r

    @staticmethod
    def from_xml(r: ET.Element) -> "Config":
        return Config(r.find("database").text, r.find("input").text, Span.from_xml(r.find("span")),
                      int(r.find("corr").text), float(np.float32(float(r.find("weight").text))),
                      r.find("normalize").text.strip().lower() == "true", int(r.find("numBreaks").text),
                      int(r.find("minSpacing").text))

    @staticmethod
    def from_xml_file(path: str) -> "Config":
        return Config.from_xml(ET.parse(path).getroot())


class ConfigBuilder:
    def __init__(self, config: Optional[Config] = None):
        self.read(config or Config())

    def read(self, c: Config):
        for k in c.__dataclass_fields__:
            setattr(self, k, getattr(c, k))

    def build(self) -> Config:
        return Config(**{k: getattr(self, k) for k in Config.__dataclass_fields__})


def native_config(c: Config, step_size: int) -> N.SegmConfig:
    return N.SegmConfig(step_size, int(c.span.has_start), int(c.span.has_stop), c.span.start or 0, c.span.stop or 0,
                        c.corrLen, c.temporalWeight, c.numBreaks, c.minSpacing)


class FeatureSegmentationImpl(ProcessorImpl):
    device = 0

    def __init__(self, config):
        super().__init__(config.build() if isinstance(config, ConfigBuilder) else config)

    def body(self) -> List[Break]:
        cfg: Config = self.config
        extr = FeatureExtractionConfig.from_xml_file(cfg.metaInput)
        norm = read_norm_file(cfg.databaseFolder, extr.num_coeffs + 1) if cfg.normalize else None
        frames, _ = read_aiff(extr.feature_output)
        self.check_aborted()
        ctx = engine.Context(self.device)
        try:
            res = engine.segm_run(ctx, native_config(cfg, extr.step_size), frames, norm)
        except N.Aborted:
            raise Aborted()
        finally:
            ctx.close()
        self.check_aborted()
        self.progress = 1.0
        return [Break(b["sim"], b["pos"]) for b in res]


class FeatureSegmentation(ProcessorFactory):
    Impl = FeatureSegmentationImpl
    Config = Config
    ConfigBuilder = ConfigBuilder
    Break = Break

    @classmethod
    def default_config(cls):
        return Config()
