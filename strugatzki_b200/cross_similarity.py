"""CrossSimilarity -- host-side mirror of CrossSimilarity.scala with a body on the B200 engine (replaces
Impl/CrossSimilarityImpl.scala:32-187).  XML tags / defaults as in
CrossSimilarity.scala:96-233.  The product is Unit; the side effect is the 1-channel float audio file."""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native as N
from . import engine
from .io import FeatureExtractionConfig, Span, read_aiff, read_norm_file, write_aiff
from .processor import Aborted, ProcessorFactory, ProcessorImpl

verbose = False


@dataclass(frozen=True)
class Config:
    databaseFolder: str = "database"
    metaInput1: str = "input1_feat.xml"
    metaInput2: str = "input2_feat.xml"
    audioOutput: str = "output.aif"
    audioOutputType: str = "aiff"          # AudioFileType.id
    span1: Span = Span.all()
    span2: Span = Span.all()
    temporalWeight: float = 0.5
    normalize: bool = True
    maxBoost: float = 8.0

    def to_xml(self) -> ET.Element:
        r = ET.Element("crosssimilarity")
        for tag, v in (("database", self.databaseFolder), ("input1", self.metaInput1), ("input2", self.metaInput2),
                       ("output", self.audioOutput), ("outputType", self.audioOutputType)):
            ET.SubElement(r, tag).text = str(v)
        for tag, sp in (("span1", self.span1), ("span2", self.span2)):
            if sp.has_start or sp.has_stop:                     # Span.All writes no element (:170-171)
                e = ET.SubElement(r, tag)
                for c in sp.to_xml():
                    e.append(c)
        ET.SubElement(r, "weight").text = repr(float(np.float32(self.temporalWeight)))
        ET.SubElement(r, "normalize").text = str(self.normalize).lower()
        ET.SubElement(r, "maxBoost").text = repr(float(np.float32(self.maxBoost)))
        return r

    @staticmethod
    def from_xml(r: ET.Element) -> "Config":
        return Config(r.find("database").text, r.find("input1").text, r.find("input2").text, r.find("output").text,
                      r.find("outputType").text, Span.from_xml(r.find("span1")), Span.from_xml(r.find("span2")),
                      float(np.float32(float(r.find("weight").text))),
                      r.find("normalize").text.strip().lower() == "true",
                      float(np.float32(float(r.find("maxBoost").text))))

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


def native_config(c: Config, step_size: int) -> N.CrossConfig:
    s1, s2 = c.span1, c.span2
    return N.CrossConfig(step_size, int(s1.has_start), int(s1.has_stop), int(s2.has_start), int(s2.has_stop), 0,
                         s1.start or 0, s1.stop or 0, s2.start or 0, s2.stop or 0, c.temporalWeight, c.maxBoost)


class CrossSimilarityImpl(ProcessorImpl):
    device = 0

    def __init__(self, config):
        super().__init__(config.build() if isinstance(config, ConfigBuilder) else config)

    def body(self) -> None:
        cfg: Config = self.config
        e1 = FeatureExtractionConfig.from_xml_file(cfg.metaInput1)
        e2 = FeatureExtractionConfig.from_xml_file(cfg.metaInput2)
        if (e1.fft_size, e1.fft_overlap, e1.num_coeffs) != (e2.fft_size, e2.fft_overlap, e2.num_coeffs):
            raise ValueError(f"requirement failed: Analysis settings for {cfg.metaInput1} and {cfg.metaInput2} differ.")
        norm = read_norm_file(cfg.databaseFolder, e1.num_coeffs + 1) if cfg.normalize else None
        f1, spec1 = read_aiff(e1.feature_output)
        f2, _ = read_aiff(e2.feature_output)
        self.check_aborted()
        ctx = engine.Context(self.device)
        try:
            sim = engine.cross_run(ctx, native_config(cfg, e1.step_size), f1, f2, norm)
        except N.Aborted:
            raise Aborted()
        finally:
            ctx.close()
        self.check_aborted()
        write_aiff(cfg.audioOutput, sim.reshape(-1, 1), spec1.sample_rate)   # 1 channel, rate of input 1 (:88-91)
        self.progress = 1.0
        return None


class CrossSimilarity(ProcessorFactory):
    Impl = CrossSimilarityImpl
    Config = Config
    ConfigBuilder = ConfigBuilder

    @classmethod
    def default_config(cls):
        return Config()
