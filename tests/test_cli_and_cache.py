"""Command line front end (Strugatzki.scala:67-440) and the on-disk database cache -- host logic, no GPU."""
import io
import os
import struct

import numpy as np
import pytest

from util import STEP, make_db, make_input
from strugatzki_b200 import cli
from strugatzki_b200.io import DatabaseCache, FeatureExtractionConfig, Span, read_aiff_many, write_aiff


def test_secs_to_frames_and_formats():
    # (s * sampleRate + 0.5).toLong, Strugatzki.scala:143
    assert cli.secs_to_frames(2.0, 44100.0) == 88200
    assert cli.secs_to_frames(0.5, 44100.0) == 22050
    assert cli.secs_to_frames(1.00001, 44100.0) == 44100          # 44100.441 + 0.5 -> 44100
    assert cli.secs_to_frames(0.00002, 44100.0) == 1              # 0.882 + 0.5 -> 1
    assert cli.secs_to_frames(0.0, 48000.0) == 0
    assert cli.to_percent_str(0.98996) == "99.0%" and cli.to_percent_str(0.5) == "50.0%"
    assert cli.to_db_str(1.0) == "0.0 dB" and cli.to_db_str(0.5) == "-6.0 dB" and cli.to_db_str(8.0) == "18.1 dB"


def test_read_sample_rate(tmp_path):
    a = str(tmp_path / "a.aif")
    write_aiff(a, np.zeros((4, 2), np.float32), 48000.0)
    assert cli.read_sample_rate(a) == 48000.0
    w = str(tmp_path / "w.wav")
    with open(w, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, 44100, 176400, 4, 16) +
                b"data" + struct.pack("<I", 0))
    assert cli.read_sample_rate(w) == 44100.0
    with pytest.raises(IOError):
        cli.read_sample_rate(__file__)


def _capture(monkeypatch):
    seen = []
    monkeypatch.setattr(cli, "_go", lambda factory, config, on_success, out: seen.append((factory, config)) or 0)
    return seen


def test_correlate_options_build_the_reference_config(monkeypatch, tmp_path):
    seen = _capture(monkeypatch)
    meta = str(tmp_path / "q_feat.xml")
    FeatureExtractionConfig(str(tmp_path / "q.aif"), str(tmp_path / "q_feat.aif"), None, 13, 1024, 2).write(meta)
    write_aiff(str(tmp_path / "q.aif"), np.zeros((8, 1), np.float32), 44100.0)      # the audio whose header gives the rate
    out = io.StringIO()
    rc = cli.main(["-c", "-d", "db", "--in-start", "0", "--in-stop", "2", "--dur-min", "1", "--dur-max", "8",
                   "--out-start", "4.0", "--out-stop", "6.0", "--out-temp", "0.25", "-m", "20", "--num-per-file", "2",
                   "--spacing", "0.5", "--boost-max", "4", "--no-norm", meta], out)
    assert rc == 0 and len(seen) == 1
    c = seen[0][1]
    assert (c.databaseFolder, c.metaInput) == ("db", meta)
    assert c.punchIn.span == Span(0, 88200) and c.punchIn.temporalWeight == 0.5
    assert c.punchOut.span == Span(176400, 264600) and c.punchOut.temporalWeight == 0.25
    assert (c.minPunch, c.maxPunch, c.numMatches, c.numPerFile, c.minSpacing) == (44100, 352800, 20, 2, 22050)
    assert c.normalize is False and c.maxBoost == 4.0
    # defaults of the reference: one match, one per file, spacing 0, boost 8, normalise
    seen.clear()
    assert cli.main(["-c", "-d", "db", "--in-start", "1", "--in-stop", "3", "--dur-min", "1", "--dur-max", "2",
                     "--sample-rate", "48000", meta], out) == 0
    c = seen[0][1]
    assert c.punchIn.span == Span(48000, 144000) and c.punchOut is None
    assert (c.numMatches, c.numPerFile, c.minSpacing, c.maxBoost, c.normalize) == (1, 1, 0, 8.0, True)
    # requirements (:146,158-162) and missing options end with exit code 1
    for bad in (["--in-start", "2", "--in-stop", "2", "--dur-min", "1", "--dur-max", "2"],
                ["--in-start", "0", "--in-stop", "2", "--dur-min", "0", "--dur-max", "2"],
                ["--in-start", "0", "--in-stop", "2", "--dur-min", "3", "--dur-max", "2"],
                ["--in-start", "0", "--in-stop", "2", "--dur-min", "1"]):
        assert cli.main(["-c", "-d", "db"] + bad + [meta], out) == 1
    assert cli.main([], out) == 1 and cli.main(["-f", "x"], out) == 1 and cli.main(["--nonsense"], out) == 1


def test_segmentation_and_selfsimilarity_options(monkeypatch, tmp_path):
    seen = _capture(monkeypatch)
    meta = str(tmp_path / "q_feat.xml")
    FeatureExtractionConfig(str(tmp_path / "q.aif"), str(tmp_path / "q_feat.aif"), None, 13, 1024, 2).write(meta)
    out = io.StringIO()
    assert cli.main(["-s", "-d", "db", "--length", "0.5", "-m", "20", "--spacing", "0.5", "--span-start", "1.5",
                     "--sample-rate", "44100", meta], out) == 0
    c = seen[-1][1]
    assert (c.corrLen, c.numBreaks, c.minSpacing, c.temporalWeight, c.normalize) == (22050, 20, 22050, 0.5, True)
    assert c.span == Span.from_(66150) and c.databaseFolder == "db"
    assert cli.main(["-s", "--sample-rate", "44100", meta], out) == 1            # normalisation needs -d (exit1)
    assert cli.main(["-s", "--no-norm", "--sample-rate", "44100", meta], out) == 0
    assert seen[-1][1].minSpacing == 8820                                         # default spacing 0.2 s
    assert cli.main(["-x", "-d", "db", "-c", "gray", "-m", "4", "-i", "--color-warp", "0.5", "--sample-rate", "44100",
                     meta, "out.png"], out) == 0
    c = seen[-1][1]
    assert (c.corrLen, c.decimation, c.colors, c.colorInv, c.colorWarp, c.imageOutput) == (44100, 4, "gray", True, 0.5, "out.png")
    assert cli.main(["-y", "--no-norm", "--span1-stop", "2", "--sample-rate", "44100", meta, meta, "o.aif"], out) == 0
    c = seen[-1][1]
    assert c.span1 == Span.until(88200) and c.span2 == Span.all() and c.audioOutput == "o.aif" and c.maxBoost == 8.0


def test_database_cache_round_trip_and_invalidation(tmp_path, monkeypatch):
    folder = str(tmp_path)
    files, _ = make_db(3, [300, 0, 450])
    paths = []
    for i, f in enumerate(files):
        p = os.path.join(folder, f"f{i}_feat.aif")
        write_aiff(p, f, 44100.0 / STEP)
        paths.append(p)
    cache = DatabaseCache(folder, paths, 14)
    assert cache.load() is None
    with cache.writer() as w:
        for frames, spec in read_aiff_many(paths, raw=True):
            w.add(frames)                                  # big-endian payloads, as the upload path sees them
    mm, counts = DatabaseCache(folder, paths, 14).load()
    assert counts == [300, 0, 450] and mm.shape == (750, 14) and mm.dtype == np.float32
    assert np.array_equal(np.asarray(mm[:300]), files[0]) and np.array_equal(np.asarray(mm[300:]), files[2])
    # a different channel count, a touched file or a different file list is a different cache
    assert DatabaseCache(folder, paths, 13).load() is None
    assert DatabaseCache(folder, paths[:2], 14).load() is None
    st = os.stat(paths[2])
    os.utime(paths[2], ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))
    assert DatabaseCache(folder, paths, 14).load() is None
    os.utime(paths[2], ns=(st.st_atime_ns, st.st_mtime_ns))
    assert DatabaseCache(folder, paths, 14).load() is not None
    # an interrupted writer publishes nothing; SGZ_DB_CACHE=0 turns the cache off
    c2 = DatabaseCache(folder, paths[:2], 14)
    with pytest.raises(RuntimeError):
        with c2.writer() as w:
            w.add(files[0])
            raise RuntimeError("decode failed")
    assert c2.load() is None and not os.path.exists(c2.base + ".f32.tmp")
    monkeypatch.setenv("SGZ_DB_CACHE", "0")
    assert DatabaseCache(folder, paths, 14).load() is None
