"""The reference-facing surface end to end: feature AIFFs + meta XML + feat_norms.aif on disk ->
FeatureCorrelation / FeatureSegmentation / SelfSimilarity processors -> results equal the oracle's."""
import os

import numpy as np
import pytest

from util import O, STEP, make_db, make_input, plant_needles, synth
from strugatzki_b200 import feature_correlation as fc
from strugatzki_b200 import feature_segmentation as fs
from strugatzki_b200 import self_similarity as ss
from strugatzki_b200 import cross_similarity as cs
from strugatzki_b200 import feature_stats as fst
from strugatzki_b200.io import FeatureExtractionConfig, Span, read_aiff, write_aiff, write_norm_file
from strugatzki_b200.processor import Progress, Result, Success

pytestmark = pytest.mark.gpu


def write_feature_file(folder, name, frames):
    feat = os.path.join(folder, f"{name}_feat.aif")
    write_aiff(feat, frames, 44100.0 / STEP)
    FeatureExtractionConfig(os.path.join(folder, f"{name}.aif"), feat, None, frames.shape[1] - 1, 1024, 2).write(
        os.path.join(folder, f"{name}_feat.xml"))
    return os.path.join(folder, f"{name}_feat.xml")


@pytest.fixture()
def database(tmp_path):
    folder = str(tmp_path / "db")
    os.makedirs(folder)
    files, norm = make_db(5, [2600, 3000, 2200, 2800, 2400])
    inp = make_input(900)
    plant_needles(files, inp[:172], [(1, 640), (3, 2000)])
    for i, f in enumerate(files):
        write_feature_file(folder, f"file{i:02d}", f)
    meta_in = write_feature_file(str(tmp_path), "query", inp)
    write_norm_file(folder, norm)
    return folder, meta_in, files, inp, norm


def test_feature_correlation_processor(ctx, database):
    folder, meta_in, files, inp, norm = database
    b = fc.ConfigBuilder()
    b.databaseFolder, b.metaInput = folder, meta_in
    b.punchIn = fc.Punch(Span(0, 88200), 0.5)
    b.numMatches, b.numPerFile, b.minSpacing = 4, 2, 22050
    events = []
    proc = fc.FeatureCorrelation.run(b.build(), events.append)
    got = proc.await_result(60)
    want = O.corr_search(O.CorrParams(step_size=STEP, input=inp, punch_in=(0, 88200), norm=norm, num_matches=4,
                                      num_per_file=2, min_spacing=22050), files)
    assert len(got) == len(want) == 4
    for g, w in zip(got, want):
        assert g.file == os.path.join(folder, f"file{w['file']:02d}.aif")
        assert (g.punch.start, g.punch.stop) == (w["start"], w["stop"])
        assert abs(g.sim - w["sim"]) <= 1e-5 * abs(w["sim"]) + 2e-6
        assert abs(g.boostIn - w["boostIn"]) <= 1e-5 * abs(w["boostIn"])
    assert (got[0].file, got[0].punch.start) in ((os.path.join(folder, "file01.aif"), 640 * STEP),
                                                 (os.path.join(folder, "file03.aif"), 2000 * STEP))
    prog = [e.amount for e in events if isinstance(e, Progress)]
    assert prog == sorted(prog) and prog[-1] == 1.0
    assert isinstance(events[-1], Result) and isinstance(events[-1].value, Success)
    # normalize = true without feat_norms.aif -> failure surfaces through the processor, like the reference
    os.remove(os.path.join(folder, "feat_norms.aif"))
    with pytest.raises(IOError):
        fc.FeatureCorrelation.run(b.build()).await_result(60)


def test_segmentation_and_selfsimilarity_processors(ctx, tmp_path):
    folder = str(tmp_path)
    f, _ = synth.regime_file(synth.BASE_SEED, 21, 1500, 14, 7)
    _, _, _, norm = synth.default_profile(14)
    meta = write_feature_file(folder, "song", f)
    write_norm_file(folder, norm)
    cfg = fs.Config(folder, meta, Span.all(), 22050, 0.5, True, 6, 22050)
    got = fs.FeatureSegmentation.run(cfg).await_result(60)
    want = O.segm_run(O.SegmParams(step_size=STEP, norm=norm, num_breaks=6), f)
    assert [(b.pos, np.float32(b.sim).tobytes()) for b in got] == \
           [(b["pos"], np.float32(b["sim"]).tobytes()) for b in want]

    png = os.path.join(folder, "out.png")
    scfg = ss.Config(folder, meta, None, png, Span.until(400 * STEP), 20480, 2, 0.5, ss.GrayScale, 1.0, 1.0, False,
                     True)
    from PIL import Image
    want_img = O.self_image(O.SelfParams(step_size=STEP, corr_len=20480, decimation=2, norm=norm,
                                         span_stop=400 * STEP), f)
    for precise in (True, False):
        ss.precise = precise
        try:
            assert ss.SelfSimilarity.run(scfg).await_result(120) is None
        finally:
            ss.precise = False
        img = np.asarray(Image.open(png).convert("RGB")).astype(np.int32)
        packed = (img[..., 0] << 16) | (img[..., 1] << 8) | img[..., 2]
        assert packed.shape == want_img.shape
        if precise:
            assert np.array_equal(packed, want_img)                       # FP64 replay: pixel identical
        else:
            assert np.abs((packed & 0xFF) - (want_img & 0xFF)).max() <= 1  # FP32 Gram: within 1 grey level
    with pytest.raises(RuntimeError):   # PsychoOptical without the third-party palette table
        ss.SelfSimilarity.run(ss.Config(folder, meta, None, png, Span.until(400 * STEP), 20480, 2, 0.5,
                                        ss.PsychoOptical)).await_result(60)


def test_cross_similarity_processor(ctx, tmp_path):
    folder = str(tmp_path)
    files, norm = make_db(2, [260, 8900])
    m1 = write_feature_file(folder, "short", files[0])
    m2 = write_feature_file(folder, "long", files[1])
    write_norm_file(folder, norm)
    out = os.path.join(folder, "cross.aif")
    cfg = cs.Config(folder, m1, m2, out, "aiff", Span.all(), Span.from_(100 * STEP), 0.5, True, 8.0)
    assert cs.Config.from_xml(cfg.to_xml()) == cfg
    assert cs.CrossSimilarity.run(cfg).await_result(60) is None
    got, spec = read_aiff(out)
    want = O.cross_run(O.CrossParams(step_size=STEP, norm=norm, span2=(100 * STEP, None)), files[0], files[1])
    assert spec.num_channels == 1 and abs(spec.sample_rate - 44100.0 / STEP) < 1e-9
    assert np.array_equal(got[:, 0].view(np.uint32), want.view(np.uint32))


def test_feature_stats_processor_writes_norm_file(ctx, tmp_path):
    folder = str(tmp_path)
    files, _ = make_db(4, [2600, 3000, 2200, 2800])
    for i, f in enumerate(files):
        write_feature_file(folder, f"file{i:02d}", f)
    res = fst.stats_of_folder(folder)
    want = O.stats_run(files)
    assert len(res) == 14 and np.allclose(np.array(res), want, rtol=1e-9)
    from strugatzki_b200.io import read_norm_file
    norm = read_norm_file(folder, 14)
    assert np.array_equal(norm.view(np.uint32), want.astype(np.float32).view(np.uint32))


def test_cli_correlate_prints_the_oracle_matches_and_reuses_the_database_cache(ctx, database, monkeypatch):
    """`Strugatzki -c` (Strugatzki.scala:101-213) through strugatzki_b200.cli; the second run maps the on-disk cache of the
    decoded database instead of parsing the AIFFs and must print the same matches"""
    import io as _io
    from strugatzki_b200 import cli
    from strugatzki_b200 import io as sio
    folder, meta_in, files, inp, norm = database
    args = ["-c", "-d", folder, "--in-start", "0", "--in-stop", "2", "--dur-min", "1", "--dur-max", "8", "-m", "4",
            "--num-per-file", "2", "--spacing", "0.5", "--sample-rate", "44100", meta_in]
    want = O.corr_search(O.CorrParams(step_size=STEP, input=inp, punch_in=(0, 88200), norm=norm, num_matches=4,
                                      num_per_file=2, min_spacing=22050), files)

    def run():
        out = _io.StringIO()
        assert cli.main(args, out) == 0
        return out.getvalue()

    first = run()
    assert "  Success." in first and first.count("#") == 25
    blocks = [b for b in first.split("\n\n") if b.strip().startswith("File")]
    assert len(blocks) == len(want) == 4
    for b, w in zip(blocks, want):
        lines = dict(l.split(None, 1) if l.startswith("File") else l.split(":", 1) for l in b.strip().split("\n"))
        assert lines["File"].strip() == os.path.join(folder, f"file{w['file']:02d}.aif")
        assert int(lines["Span start"]) == w["start"]
        assert lines["Similarity"].strip() == cli.to_percent_str(w["sim"])
        assert lines["Boost in  "].strip() == cli.to_db_str(w["boostIn"])
    assert any(n.endswith(".f32") for n in os.listdir(os.path.join(folder, ".sgz_cache")))
    calls = []
    real = sio.read_aiff_many
    monkeypatch.setattr("strugatzki_b200.feature_correlation.read_aiff_many", lambda *a, **k: calls.append(1) or real(*a, **k))
    assert run() == first and calls == []                      # served from the cache, identical output
