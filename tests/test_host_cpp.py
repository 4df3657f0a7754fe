"""The compiled-language host layer (host/strugatzki_host.hpp, C++17 over the C ABI): the reference's own XML
round-trip test on CPU, and on the GPU the same searches as the Python mirror, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from util import O, ROOT, STEP, make_db, make_input, plant_needles, synth
from strugatzki_b200 import _native as N
from strugatzki_b200.io import FeatureExtractionConfig, write_aiff, write_norm_file


@pytest.fixture(scope="module")
def host_bin(tmp_path_factory):
    N.build()
    out = str(tmp_path_factory.mktemp("hostbin") / "host_test")
    lib_dir = os.path.join(ROOT, "strugatzki_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(ROOT, "host"), os.path.join(ROOT, "tests", "cpp", "host_test.cpp"), "-o", out,
                           "-L", lib_dir, "-lsgz_b200", f"-Wl,-rpath,{lib_dir}", "-lpthread"])
    return out


def write_feature_file(folder, name, frames):
    feat = os.path.join(folder, f"{name}_feat.aif")
    write_aiff(feat, frames, 44100.0 / STEP)
    FeatureExtractionConfig(os.path.join(folder, f"{name}.aif"), feat, None, frames.shape[1] - 1, 1024, 2).write(
        os.path.join(folder, f"{name}_feat.xml"))
    return os.path.join(folder, f"{name}_feat.xml")


def make_folder(tmp_path):
    folder = str(tmp_path / "db")
    os.makedirs(folder)
    files, norm = make_db(5, [2600, 3000, 2200, 2800, 2400])
    inp = make_input(900)
    plant_needles(files, inp[:172], [(1, 640), (3, 2000)])
    files[3][2400:2572] = synth.plant(inp[345:517], 5, 9)
    for i, f in enumerate(files):
        write_feature_file(folder, f"file{i:02d}", f)
    meta_in = write_feature_file(str(tmp_path), "query", inp)
    write_norm_file(folder, norm)
    return folder, meta_in, files, inp, norm


def test_cpp_host_xml_roundtrip(host_bin):
    out = subprocess.run([host_bin, "xml"], capture_output=True, text=True)
    assert out.returncode == 0 and "xml ok" in out.stdout, out.stdout + out.stderr


def test_cpp_host_fails_loudly_without_gpu(host_bin, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    folder, meta_in, *_ = make_folder(tmp_path)
    out = subprocess.run([host_bin, "corr", folder, meta_in, "3", "1", "0"], capture_output=True, text=True)
    assert out.returncode == 4 and "no CPU fallback" in out.stdout, out.stdout


@pytest.mark.gpu
def test_cpp_host_processors_match_oracle(host_bin, tmp_path):
    folder, meta_in, files, inp, norm = make_folder(tmp_path)

    def parse(stdout, key):
        return [l.split()[1:] for l in stdout.splitlines() if l.startswith(key + " ")]

    for po in (None, (345 * STEP, 517 * STEP)):
        args = [host_bin, "corr", folder, meta_in, "4", "2", "22050"]
        if po:
            args += [str(po[0]), str(po[1]), str(86 * STEP), str(689 * STEP)]
        out = subprocess.run(args, capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        got = parse(out.stdout, "match")
        want = O.corr_search(O.CorrParams(step_size=STEP, input=inp, punch_in=(0, 88200), punch_out=po,
                                          min_punch=86 * STEP if po else 22050, max_punch=689 * STEP if po else 88200,
                                          norm=norm, num_matches=4, num_per_file=2, min_spacing=22050), files)
        assert len(got) == len(want) == 4
        for g, w in zip(got, want):
            assert g[1] == os.path.join(folder, f"file{w['file']:02d}.aif")
            assert (int(g[2]), int(g[3])) == (w["start"], w["stop"])
            assert abs(float(g[0]) - w["sim"]) <= 1e-5 * abs(w["sim"]) + 2e-6
        ev = parse(out.stdout, "events")[0]
        assert int(ev[0]) >= 1 and int(ev[1]) == 1                      # Progress events, one Result

    seg, _ = synth.regime_file(synth.BASE_SEED, 21, 1500, 14, 7)
    meta_s = write_feature_file(str(tmp_path), "song", seg)
    out = subprocess.run([host_bin, "segm", folder, meta_s, "6"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    wb = O.segm_run(O.SegmParams(step_size=STEP, norm=norm, num_breaks=6), seg)
    gb = parse(out.stdout, "break")
    assert [(int(b[1]), np.float32(float(b[0])).tobytes()) for b in gb] == \
           [(b["pos"], np.float32(b["sim"]).tobytes()) for b in wb]

    png = str(tmp_path / "self.png")
    want_img = O.self_image(O.SelfParams(step_size=STEP, corr_len=20480, decimation=2, norm=norm), seg)
    from PIL import Image
    for precise in (1, 0):
        out = subprocess.run([host_bin, "self", folder, meta_s, png, "2", str(precise)], capture_output=True, text=True)
        assert out.returncode == 0 and f"imgExt {want_img.shape[0]}" in out.stdout, out.stdout
        img = np.asarray(Image.open(png).convert("RGB")).astype(np.int32)       # also validates the PNG writer
        packed = (img[..., 0] << 16) | (img[..., 1] << 8) | img[..., 2]
        if precise:
            assert np.array_equal(packed, want_img)
        else:
            assert np.abs((packed & 0xFF) - (want_img & 0xFF)).max() <= 1
