"""Host-side mirror of the reference's Config / Processor surface.  The first test is the reference's own
(and only) test, StrugatzkiSuite.scala:12-92: mutate every field, round-trip through XML, compare."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from strugatzki_b200 import feature_correlation as fc
from strugatzki_b200 import feature_segmentation as fs
from strugatzki_b200 import self_similarity as ss
from strugatzki_b200.io import (FeatureExtractionConfig, Span, full_to_feat, list_database, read_aiff,
                                read_norm_file, span_spacing, write_aiff, write_norm_file)
from strugatzki_b200.processor import Aborted, Failure, Progress, ProcessorImpl, Result, Success


def rt(x):
    return type(x).from_xml(ET.fromstring(ET.tostring(x.to_xml())))


def test_xml_roundtrip_like_StrugatzkiSuite():
    fe1 = FeatureExtractionConfig(os.path.abspath("testing.aif"), "relative.aif", None, 14, 1025, 3)
    fe2 = FeatureExtractionConfig(fe1.audio_input, "relative.aif", os.path.dirname(fe1.audio_input), 14, 1025, 3)
    assert rt(fe1) == fe1 and rt(fe2) == fe2

    b = fc.ConfigBuilder()
    b.databaseFolder = os.path.abspath("db")
    b.metaInput = "rarara.xml"
    old = b.punchIn
    b.punchIn = fc.Punch(Span(old.span.start + 1, old.span.stop + 2), float(np.float32(old.temporalWeight + 0.11)))
    b.punchOut = fc.Punch(Span(555, 666), float(np.float32(0.1234)))
    b.minPunch += 1; b.maxPunch += 2; b.normalize = not b.normalize; b.maxBoost += 1
    b.numMatches += 1; b.numPerFile += 1; b.minSpacing += 1
    c1 = b.build()
    b.punchOut = None
    b.normalize = not b.normalize
    c2 = b.build()
    assert rt(c1) == c1 and rt(c2) == c2 and c1 != c2

    m1 = fc.Match(float(np.float32(0.23)), "gaga.aif", Span(33, 44), -6.0, -7.0)
    m2 = fc.Match(float(np.float32(0.46)), os.path.abspath("rara.wav"), Span(666, 777), 8.0, 9.0)
    assert rt(m1) == m1 and rt(m2) == m2

    s = fs.ConfigBuilder()
    s.databaseFolder = os.path.abspath("db"); s.metaInput = "rarara.xml"; s.span = Span(1, 2)
    s.corrLen += 1; s.temporalWeight = float(np.float32(s.temporalWeight + 0.1)); s.normalize = not s.normalize
    s.numBreaks += 1; s.minSpacing += 1
    s1 = s.build()
    s.span = Span.all(); s.normalize = not s.normalize
    s2 = s.build()
    assert rt(s1) == s1 and rt(s2) == s2
    assert rt(fs.Break(float(np.float32(0.5)), 12345)) == fs.Break(0.5, 12345)

    # not covered upstream: SelfSimilarity.Config
    x = ss.ConfigBuilder()
    x.metaInput2 = "other_feat.xml"; x.span = Span.from_(1000); x.decimation = 3; x.colors = ss.GrayScale
    x.colorWarp = 0.5; x.colorCeil = 0.75; x.colorInv = True; x.normalize = False
    assert rt(x.build()) == x.build()
    assert rt(ss.Config()) == ss.Config()
    with pytest.raises(ValueError):
        ss.color_scheme("rainbow")


def test_defaults_are_the_references():
    c = fc.Config()
    assert (c.databaseFolder, c.metaInput, c.punchIn, c.punchOut, c.minPunch, c.maxPunch, c.normalize, c.maxBoost,
            c.numMatches, c.numPerFile, c.minSpacing) == \
        ("database", "input_feat.xml", fc.Punch(Span(0, 44100), 0.5), None, 22050, 88200, True, 8.0, 1, 1, 0)
    s = fs.Config()
    assert (s.span, s.corrLen, s.temporalWeight, s.normalize, s.numBreaks, s.minSpacing) == \
        (Span.all(), 22050, 0.5, True, 1, 22050)
    x = ss.Config()
    assert (x.imageOutput, x.corrLen, x.decimation, x.colors, x.colorWarp, x.colorCeil, x.colorInv) == \
        ("output_selfsim.png", 44100, 1, ss.PsychoOptical, 1.0, 1.0, False)
    e = FeatureExtractionConfig()
    assert (e.num_coeffs, e.fft_size, e.fft_overlap, e.step_size) == (13, 1024, 2, 512)


def test_span_and_rounding_helpers():
    assert span_spacing(Span(0, 10), Span(15, 20)) == 5 and span_spacing(Span(15, 20), Span(0, 10)) == 5
    assert span_spacing(Span(0, 10), Span(5, 20)) == -5                     # overlapping -> negative
    assert full_to_feat(88200, 512) == 172 and full_to_feat(22050, 512) == 43 and full_to_feat(44100, 512) == 86
    assert full_to_feat(255, 512) == 0 and full_to_feat(256, 512) == 1
    assert Span.from_xml(None) == Span.all() and Span.until(7).has_stop and not Span.until(7).has_start


def test_aiff_roundtrip_and_norm_file(tmp_path):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((333, 14)).astype(np.float32)
    p = str(tmp_path / "x_feat.aif")
    write_aiff(p, a, 86.1328125)
    b, spec = read_aiff(p)
    assert np.array_equal(a, b) and (spec.num_channels, spec.num_frames) == (14, 333) and spec.big_endian_f32
    assert abs(spec.sample_rate - 86.1328125) < 1e-9
    raw, _ = read_aiff(p, raw=True)
    assert raw.dtype == np.dtype(">f4") and np.array_equal(raw.astype(np.float32), a)
    norm = np.stack([np.arange(14), np.arange(14) + 2.5], 1).astype(np.float32)
    write_norm_file(str(tmp_path), norm)
    assert np.array_equal(read_norm_file(str(tmp_path), 14), norm)
    with pytest.raises(ValueError):
        read_norm_file(str(tmp_path), 13)                                   # reference: require(...)
    with pytest.raises(IOError):
        (tmp_path / "bad.aif").write_bytes(b"RIFFxxxxWAVE")
        read_aiff(str(tmp_path / "bad.aif"))


def test_database_discovery(tmp_path):
    d = tmp_path / "db"
    d.mkdir()
    for name, nc, fft in (("b", 13, 1024), ("a", 13, 1024), ("c", 12, 1024), ("d", 13, 2048), ("in", 13, 1024)):
        FeatureExtractionConfig(f"{name}.aif", str(d / f"{name}_feat.aif"), None, nc, fft, 2).write(
            str(d / f"{name}_feat.xml"))
    (d / "notes.txt").write_text("x")
    got = list_database(str(d), str(d / "in_feat.xml"), 13, 512)
    assert [e.audio_input for e in got] == ["a.aif", "b.aif"]              # same numCoeffs + step, minus metaInput


def test_processor_contract():
    events = []

    class P(ProcessorImpl):
        def body(self):
            for i in range(5):
                self.check_aborted()
                self.progress = (i + 1) / 5
            return 42

    p = P(None)
    p.add_listener(events.append)
    p.start()
    assert p.await_result(5) == 42 and p.is_completed
    assert [e.amount for e in events if isinstance(e, Progress)] == [0.2, 0.4, 0.6, 0.8, 1.0]
    assert isinstance(events[-1], Result) and isinstance(events[-1].value, Success)

    class Q(ProcessorImpl):
        def body(self):
            import time
            while True:
                self.check_aborted()
                time.sleep(0.001)

    q = Q(None)
    res = []
    q.add_listener(res.append)
    q.start()
    q.abort()
    with pytest.raises(Aborted):
        q.await_result(5)
    assert isinstance(res[-1].value, Failure) and isinstance(res[-1].value.exception, Aborted)


def test_read_aiff_many_keeps_order_and_reports_errors(tmp_path):
    """ingest pipeline: files come back in the order of the list (it fixes the file indices of a search) whatever the
    readers' completion order; a broken file raises at its position; closing the generator early cancels the rest"""
    from strugatzki_b200 import io as sio
    rng = np.random.default_rng(0)
    paths, datas = [], []
    for i in range(23):
        n = int(rng.integers(1, 400)) if i != 7 else 20000      # one big file finishes late
        d = rng.standard_normal((n, 3)).astype(np.float32)
        p = str(tmp_path / f"f{i:02d}_feat.aif")
        sio.write_aiff(p, d)
        paths.append(p)
        datas.append(d)
    got = list(sio.read_aiff_many(paths, raw=False, workers=4, window=6))
    assert len(got) == len(paths)
    for (frames, spec), d in zip(got, datas):
        assert spec.num_channels == 3 and np.array_equal(np.asarray(frames, np.float32), d)
    raw = list(sio.read_aiff_many(paths[:3], raw=True, workers=2))
    assert all(np.array_equal(np.asarray(f, np.float32), d) for (f, _), d in zip(raw, datas))
    bad = paths[:5] + [str(tmp_path / "missing.aif")] + paths[5:]
    seen = 0
    with pytest.raises(Exception):
        for _ in sio.read_aiff_many(bad, workers=3, window=4):
            seen += 1
    assert seen == 5
    gen = sio.read_aiff_many(paths, workers=4, window=4)
    next(gen)
    gen.close()          # early abort: no hang, readers shut down
