"""SelfSimilarity -- host-side mirror of Api/SelfSimilarity.scala with a body on the B200 engine (replaces
Impl/SelfSimilarityImpl.scala:31-180).  The engine returns the BufferedImage.TYPE_INT_RGB pixel array; the
PNG container is written here (`ImageIO.write`, :167).

`PsychoOptical` needs de.sciss:intensitypalette's table, which is a third-party artefact that is not in the
reference tree; supply it with `set_intensity_palette(lut)` (int RGB per index) -- without it the scheme
raises, like a missing dependency would.  `GrayScale` is fully specified in-source (:99-107).
"""
from __future__ import annotations

import struct
import xml.etree.ElementTree as ET
import zlib
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native as N
from . import engine
from .io import FeatureExtractionConfig, Span, read_aiff, read_norm_file
from .processor import Aborted, ProcessorFactory, ProcessorImpl

verbose = False
# Engine knob without counterpart in the reference Config: False = tiled FP32 Gram kernel (sims within 1e-5 of
# the reference, grey levels within 1 LSB), True = per-cell FP64 replay (pixel-identical image, ~30x slower).
precise = False

GrayScale = "gray"
PsychoOptical = "psycho"
_palette: Optional[np.ndarray] = None


def set_intensity_palette(lut) -> None:
    global _palette
    _palette = np.ascontiguousarray(lut, np.int32)


def color_scheme(logical_name: str) -> str:      # ColorScheme.apply, :31-34
    if logical_name in (GrayScale, PsychoOptical):
        return logical_name
    raise ValueError(f"MatchError: {logical_name}")


@dataclass(frozen=True)
class Config:
    """Api/SelfSimilarity.scala:153-240 (defaults) and :61-143 (fields)."""
    databaseFolder: str = "database"
    metaInput: str = "input_feat.xml"
    metaInput2: Optional[str] = None
    imageOutput: str = "output_selfsim.png"
    span: Span = Span.all()
    corrLen: int = 44100
    decimation: int = 1
    temporalWeight: float = 0.5
    colors: str = PsychoOptical
    colorWarp: float = 1.0
    colorCeil: float = 1.0
    colorInv: bool = False
    normalize: bool = True

    def to_xml(self) -> ET.Element:
        r = ET.Element("selfsimilarity")
        ET.SubElement(r, "database").text = self.databaseFolder
        ET.SubElement(r, "input").text = self.metaInput
        if self.metaInput2 is not None:
            ET.SubElement(r, "input2").text = self.metaInput2
        ET.SubElement(r, "output").text = self.imageOutput
        if self.span.has_start or self.span.has_stop:
            sp = ET.SubElement(r, "span")
            for c in self.span.to_xml():
                sp.append(c)
        for tag, v in (("corr", self.corrLen), ("decimation", self.decimation),
                       ("weight", repr(float(np.float32(self.temporalWeight)))), ("colors", self.colors),
                       ("colorWarp", repr(float(np.float32(self.colorWarp)))),
                       ("colorCeil", repr(float(np.float32(self.colorCeil)))),
                       ("colorInv", str(self.colorInv).lower()), ("normalize", str(self.normalize).lower())):
            ET.SubElement(r, tag).text = str(v)
        return r

    @staticmethod
    def from_xml(r: ET.Element) -> "Config":
        i2 = r.find("input2")
        f32 = lambda t: float(np.float32(float(r.find(t).text)))  # noqa: E731
        return Config(r.find("database").text, r.find("input").text, None if i2 is None else i2.text,
                      r.find("output").text, Span.from_xml(r.find("span")), int(r.find("corr").text),
                      int(r.find("decimation").text), f32("weight"), color_scheme(r.find("colors").text),
                      f32("colorWarp"), f32("colorCeil"), r.find("colorInv").text.strip().lower() == "true",
                      r.find("normalize").text.strip().lower() == "true")

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


def write_png(path: str, rgb: np.ndarray) -> None:
    """8-bit RGB PNG from packed 0x00RRGGBB int32 pixels (row-major)."""
    h, w = rgb.shape
    px = np.empty((h, w, 3), np.uint8)
    px[..., 0] = (rgb >> 16) & 0xFF
    px[..., 1] = (rgb >> 8) & 0xFF
    px[..., 2] = rgb & 0xFF
    try:
        from PIL import Image
        Image.fromarray(px, "RGB").save(path, format="PNG")
        return
    except ImportError:
        pass
    raw = b"".join(b"\0" + px[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def native_config(c: Config, step_size: int, lut: Optional[np.ndarray]) -> N.SelfConfig:
    return N.SelfConfig(step_size, int(c.span.has_start), int(c.span.has_stop), c.span.start or 0, c.span.stop or 0,
                        c.corrLen, c.decimation, c.temporalWeight, int(c.colorInv), c.colorWarp, c.colorCeil,
                        None if lut is None else lut.ctypes.data, 0 if lut is None else int(lut.shape[0]),
                        int(bool(precise)))


class SelfSimilarityImpl(ProcessorImpl):
    device = 0

    def __init__(self, config):
        super().__init__(config.build() if isinstance(config, ConfigBuilder) else config)

    def body(self) -> None:
        cfg: Config = self.config
        extr1 = FeatureExtractionConfig.from_xml_file(cfg.metaInput)
        extr2 = extr1 if cfg.metaInput2 is None else FeatureExtractionConfig.from_xml_file(cfg.metaInput2)
        if not (extr1.fft_size == extr2.fft_size and extr1.fft_overlap == extr2.fft_overlap
                and extr1.num_coeffs == extr2.num_coeffs):
            raise ValueError("requirement failed")                              # :34-35
        if cfg.decimation < 1:
            raise ValueError(f"requirement failed: Illegal decimation setting of {cfg.decimation}")
        lut = None
        if cfg.colors == PsychoOptical:
            if _palette is None:
                raise RuntimeError("PsychoOptical needs de.sciss.intensitypalette's table: call "
                                   "self_similarity.set_intensity_palette(lut) or use GrayScale")
            lut = _palette
        norm = read_norm_file(cfg.databaseFolder, extr1.num_coeffs + 1) if cfg.normalize else None
        f1, _ = read_aiff(extr1.feature_output)
        f2 = None
        if extr1.feature_output != extr2.feature_output:                        # :62
            f2, _ = read_aiff(extr2.feature_output)
        self.check_aborted()
        ctx = engine.Context(self.device)
        try:
            rgb, geom = engine.self_run(ctx, native_config(cfg, extr1.step_size, lut), f1, f2, norm)
        except N.Aborted:
            raise Aborted()
        finally:
            ctx.close()
        if verbose:
            print(f"Image extent is {geom['imgExt']} (yielding a matrix of {geom['imgExt'] ** 2} pixels)")
        self.check_aborted()
        write_png(cfg.imageOutput, rgb)                                         # ImageIO.write, :167
        self.progress = 1.0
        return None


class SelfSimilarity(ProcessorFactory):
    Impl = SelfSimilarityImpl
    Config = Config
    ConfigBuilder = ConfigBuilder

    @classmethod
    def default_config(cls):
        return Config()
