"""Generates tests/golden/golden.{npz,json}: outputs of the CPU oracle on seeded synthetic inputs.

The reference ships no golden vectors and cannot run here (Scala/JVM, no JDK in the image), so these
fixtures pin the ORACLE (not the reference): run `python tests/golden/make_golden.py` to regenerate after a
deliberate oracle change."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from strugatzki_b200 import synth  # noqa: E402

STEP = 512


def _matches(ms):
    return {"sim": np.array([m["sim"] for m in ms], np.float32),
            "pos": np.array([[m["file"], m["start"], m["stop"]] for m in ms], np.int64).reshape(-1, 3),
            "boost": np.array([[m["boostIn"], m["boostOut"]] for m in ms], np.float32).reshape(-1, 2)}


def cases():
    mu, sigma, floor0, norm = synth.default_profile(14)

    def db(n, frames):
        return [synth.synth_file(synth.BASE_SEED, 1 + i, frames, mu, sigma, floor0) for i in range(n)]

    inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)

    def corr_in():
        files = db(6, 2400)
        files[2][700:872] = synth.plant(inp[:172], 5, 1)
        files[4][10:182] = synth.plant(inp[:172], 5, 2)
        p = O.CorrParams(step_size=STEP, input=inp, punch_in=(0, 88200), norm=norm, num_matches=5, num_per_file=2,
                         min_spacing=22050)
        out = _matches(O.corr_search(p, files))
        out["curve"], out["boostcurve"] = O.corr_curve(p, files[2])
        return out

    def corr_inout():
        files = db(4, 2600)
        files[1][400:572] = synth.plant(inp[:172], 6, 1)
        files[1][900:1072] = synth.plant(inp[345:517], 6, 2)
        p = O.CorrParams(step_size=STEP, input=inp, punch_in=(0, 88200), punch_out=(176400, 264600), norm=norm,
                         min_punch=44100, max_punch=352800, num_matches=4, num_per_file=2, min_spacing=22050)
        return _matches(O.corr_search(p, files))

    def segm():
        f, _ = synth.regime_file(synth.BASE_SEED, 11, 3000, 14, 10)
        br, curve = O.segm_run(O.SegmParams(step_size=STEP, norm=norm, num_breaks=8), f, want_curve=True)
        return {"sim": np.array([b["sim"] for b in br], np.float32), "pos": np.array([b["pos"] for b in br], np.int64),
                "curve": curve[:3000 - 86 + 1]}

    def selfsim():
        f, _ = synth.regime_file(synth.BASE_SEED, 12, 300, 14, 4)
        img = O.self_image(O.SelfParams(step_size=STEP, corr_len=10240, decimation=2, norm=norm, color_warp=0.7), f)
        return {"img": img.astype(np.int32)}

    return {"corr_in": corr_in, "corr_inout": corr_inout, "segm": segm, "selfsim": selfsim}


if __name__ == "__main__":
    arrays = {}
    for name, fn in cases().items():
        for key, val in fn().items():
            arrays[f"{name}.{key}"] = np.asarray(val)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    json.dump({"cases": sorted(cases().keys()), "keys": sorted(arrays.keys()),
               "generator": "tests/golden/make_golden.py", "source": "oracle/sgz_oracle.c (parity unpinned)"},
              open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", len(arrays), "arrays")
