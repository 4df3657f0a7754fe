"""Command line front end with the reference's options (Strugatzki.scala:67-440): `python -m strugatzki_b200 -c|-s|-x|-y|--stats ...`.

Same option names, defaults, `secsToFrames(s) = (s * sampleRate + 0.5).toLong` rounding (:143) and result printing
(:177-211) as `Strugatzki.main`; the work runs on the B200 engine through the host mirror of the Config / Processor API.
`-f` (feature extraction) needs SuperCollider and is outside this engine.  One addition: `--sample-rate`, for when the
original audio file named in the meta file (whose header the reference reads for the rate) is not at hand.
"""
from __future__ import annotations

import argparse
import math
import os
import struct
import sys
from typing import List, Optional

from . import cross_similarity as cs
from . import feature_correlation as fc
from . import feature_segmentation as fs
from . import feature_stats as fst
from . import self_similarity as ss
from .io import FeatureExtractionConfig, Span, _from_ext80
from .processor import Aborted, Failure, Progress, Result, Success

NAME = "Strugatzki"


def secs_to_frames(s: float, sample_rate: float) -> int:
    """(s * inSpec.sampleRate + 0.5).toLong -- truncation toward zero like the JVM's d2l"""
    return int(s * sample_rate + 0.5)


def read_sample_rate(audio_path: str) -> float:
    """AudioFile.readSpec(metaIn.audioInput).sampleRate for AIFF / AIFF-C / WAVE headers"""
    with open(audio_path, "rb") as f:
        head = f.read(1 << 16)
    if head[:4] == b"FORM" and head[8:12] in (b"AIFF", b"AIFC"):
        pos = 12
        while pos + 8 <= len(head):
            cid, size = head[pos:pos + 4], struct.unpack(">I", head[pos + 4:pos + 8])[0]
            if cid == b"COMM":
                return _from_ext80(head[pos + 16:pos + 26])
            pos += 8 + size + (size & 1)
    if head[:4] == b"RIFF" and head[8:12] == b"WAVE":
        pos = 12
        while pos + 8 <= len(head):
            cid, size = head[pos:pos + 4], struct.unpack("<I", head[pos + 4:pos + 8])[0]
            if cid == b"fmt ":
                return float(struct.unpack("<I", head[pos + 12:pos + 16])[0])
            pos += 8 + size + (size & 1)
    raise IOError(f"{audio_path}: cannot read the sample rate (AIFF or WAVE header expected)")


def to_percent_str(d: float) -> str:       # percentFormat: one fraction digit, no grouping
    return "NaN" if d != d else f"{d * 100:.1f}%"


def to_db_str(amp: float) -> str:          # decibelFormat of 20 log10(amp)
    if amp != amp or amp < 0:
        return "NaN"
    return ("-∞ dB" if amp == 0 else f"{20 * math.log10(amp):.1f} dB")


class _Exit(Exception):
    def __init__(self, code):
        self.code = code


def _parser(prog: str) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog=prog, add_help=True)
    p.add_argument("-v", "--verbose", action="store_true", help="Verbose output")
    p.add_argument("--sample-rate", type=float, default=None,
                   help="Sample rate of the original audio (default: read from the audio file named in the meta file)")
    p.add_argument("--device", type=int, default=0, help="CUDA device index (default 0)")
    return p


def _rate(args, meta_path: str) -> float:
    if args.sample_rate:
        return args.sample_rate
    return read_sample_rate(FeatureExtractionConfig.from_xml_file(meta_path).audio_input)


def _span(start: Optional[float], stop: Optional[float], sr: float) -> Span:
    if start is not None and stop is not None:
        return Span(secs_to_frames(start, sr), secs_to_frames(stop, sr))
    if start is not None:
        return Span.from_(secs_to_frames(start, sr))
    if stop is not None:
        return Span.until(secs_to_frames(stop, sr))
    return Span.all()


def _non_empty(span: Span) -> bool:       # SpanLike.nonEmpty: open-ended spans are never empty
    return not (span.has_start and span.has_stop and span.stop <= span.start)


def _require(cond: bool, msg: str):
    if not cond:
        raise ValueError(f"requirement failed: {msg}")     # IllegalArgumentException of `require`


def _go(factory, config, on_success, out):
    """go(factory)(config)(observer): run, print 25 progress hashes, report like Strugatzki.scala:177-211"""
    state = {"prog": 0, "code": 0}

    def observer(msg):
        if isinstance(msg, Progress):
            i = int(msg.amount * 25)
            while state["prog"] < i:
                out.write("#")
                out.flush()
                state["prog"] += 1
        elif isinstance(msg, Result):
            v = msg.value
            if isinstance(v, Success):
                on_success(v.value)
            elif isinstance(v, Failure) and isinstance(v.exception, Aborted):
                out.write("  Aborted\n")
                state["code"] = 1
            else:
                out.write("  Failed: \n")
                out.write(f"{v.exception!r}\n")
                state["code"] = 1

    proc = factory.run(config, observer)
    try:
        proc.await_result()
    except Exception:
        pass            # reported through the observer
    return state["code"]


def feature_corr(argv: List[str], out=sys.stdout) -> int:
    p = _parser(f"{NAME} -c")
    p.add_argument("-d", "--dir", required=True, help="Database directory")
    p.add_argument("--in-start", type=float, required=True, help="Punch in begin (secs)")
    p.add_argument("--in-stop", type=float, required=True, help="Punch in end (secs)")
    p.add_argument("--in-temp", type=float, default=0.5, help="Temporal weight for punch in (0 to 1, default 0.5)")
    p.add_argument("--out-start", type=float, default=None, help="Punch out begin (secs)")
    p.add_argument("--out-stop", type=float, default=None, help="Punch out end (secs)")
    p.add_argument("--out-temp", type=float, default=0.5, help="Temporal weight for punch out (0 to 1, default 0.5)")
    p.add_argument("--dur-min", type=float, required=True, help="Minimum fill duration (secs)")
    p.add_argument("--dur-max", type=float, required=True, help="Maximum fill duration (secs)")
    p.add_argument("--boost-max", type=float, default=8.0, help="Maximum loudness boost factor (default 8)")
    p.add_argument("-m", "--num-matches", type=int, default=1, help="Maximum number of matches (default 1)")
    p.add_argument("--num-per-file", type=int, default=1, help="Maximum matches per single file (default 1)")
    p.add_argument("--spacing", type=float, default=0.0, help="Minimum spacing between matches within one file (default 0.0)")
    p.add_argument("--no-norm", action="store_true", help="Do not apply feature normalization")
    p.add_argument("input", help="Meta file of input to process")
    a = p.parse_args(argv)
    sr = _rate(a, a.input)
    if (a.out_start is None) != (a.out_stop is None):
        return 0            # the reference silently does nothing when only one of the two is given (:145-153)
    punch_out = None
    if a.out_start is not None:
        span = Span(secs_to_frames(a.out_start, sr), secs_to_frames(a.out_stop, sr))
        _require(span.length > 0, "Punch out span is empty")
        punch_out = fc.Punch(span, a.out_temp)
    in_span = Span(secs_to_frames(a.in_start, sr), secs_to_frames(a.in_stop, sr))
    _require(in_span.length > 0, "Punch in span is empty")
    min_frames = secs_to_frames(a.dur_min, sr)
    _require(min_frames > 0, "Minimum duration is zero")
    max_frames = secs_to_frames(a.dur_max, sr)
    _require(max_frames >= min_frames, "Maximum duration is smaller than minimum duration")
    fc.verbose = a.verbose
    b = fc.ConfigBuilder()
    b.databaseFolder, b.metaInput = a.dir, a.input
    b.punchIn, b.punchOut = fc.Punch(in_span, a.in_temp), punch_out
    b.minPunch, b.maxPunch, b.normalize, b.maxBoost = min_frames, max_frames, not a.no_norm, a.boost_max
    b.numMatches, b.numPerFile, b.minSpacing = a.num_matches, a.num_per_file, secs_to_frames(a.spacing, sr)

    def success(res):
        if not res:
            out.write("  No matches found.\n")
            return
        out.write("  Success.\n")
        for m in res:
            out.write(f"\nFile      {os.path.abspath(m.file)}\nSimilarity: {to_percent_str(m.sim)}\n"
                      f"Span start: {m.punch.start}\nBoost in  : {to_db_str(m.boostIn)}\n")
            if punch_out is not None:
                out.write(f"Span stop : {m.punch.stop}\nBoost out : {to_db_str(m.boostOut)}\n")
        out.write("\n")

    fc.FeatureCorrelationImpl.device = a.device
    return _go(fc.FeatureCorrelation, b.build(), success, out)


def feature_segm(argv: List[str], out=sys.stdout) -> int:
    p = _parser(f"{NAME} -s")
    p.add_argument("-d", "--dir", default=None, help="Database directory (required for normalization file)")
    p.add_argument("--length", type=float, default=0.5, help="Correlation length in secs (default: 0.5)")
    p.add_argument("--temp", type=float, default=0.5, help="Temporal weight (0 to 1, default 0.5)")
    p.add_argument("--span-start", type=float, default=None, help="Search begin in file (secs)")
    p.add_argument("--span-stop", type=float, default=None, help="Search end in file (secs)")
    p.add_argument("-m", "--num-breaks", type=int, default=1, help="Maximum number of breaks (default 1)")
    p.add_argument("--spacing", type=float, default=0.2, help="Minimum spacing between matches within one file (default 0.2)")
    p.add_argument("--no-norm", action="store_true", help="Do not apply feature normalization")
    p.add_argument("input", help="Meta file of input to process")
    a = p.parse_args(argv)
    sr = _rate(a, a.input)
    span = _span(a.span_start, a.span_stop, sr)
    _require(_non_empty(span), "Span is empty")
    corr = secs_to_frames(a.length, sr)
    _require(corr > 0, "Correlation duration is zero")
    if not a.no_norm and a.dir is None:
        p.print_usage(out)
        return 1
    fs.verbose = a.verbose
    b = fs.ConfigBuilder()
    b.metaInput, b.span, b.corrLen, b.temporalWeight = a.input, span, corr, a.temp
    b.normalize, b.numBreaks, b.minSpacing = not a.no_norm, a.num_breaks, secs_to_frames(a.spacing, sr)
    if a.dir is not None:
        b.databaseFolder = a.dir

    def success(res):
        if not res:
            out.write("  No breaks found.\n")
            return
        out.write("  Success.\n")
        for br in res:
            out.write(f"\nSimilarity: {to_percent_str(br.sim)}\nPosition:   {br.pos}\n")
        out.write("\n")

    return _go(fs.FeatureSegmentation, b.build(), success, out)


def feature_self(argv: List[str], out=sys.stdout) -> int:
    p = _parser(f"{NAME} -x")
    p.add_argument("-d", "--dir", default=None, help="Database directory (required for normalization file)")
    p.add_argument("--length", type=float, default=1.0, help="Correlation length in secs (default: 1.0)")
    p.add_argument("--temp", type=float, default=0.5, help="Temporal weight (0 to 1, default 0.5)")
    p.add_argument("--span-start", type=float, default=None, help="Correlation begin in file (secs)")
    p.add_argument("--span-stop", type=float, default=None, help="Correlation end in file (secs)")
    p.add_argument("-c", "--colors", default="psycho", help="Color scale (gray|psycho ; defaults to 'psycho')")
    p.add_argument("--color-warp", type=float, default=1.0, help="Color scale warping factor (default: 1.0)")
    p.add_argument("--color-ceil", type=float, default=1.0, help="Color scale input ceiling (default: 1.0)")
    p.add_argument("-i", "--color-inv", action="store_true", help="Inverted color scale")
    p.add_argument("-m", "--decim", type=int, default=1, help="Pixel decimation factor (default: 1)")
    p.add_argument("--no-norm", action="store_true", help="Do not apply feature normalization")
    p.add_argument("--input2", default=None, help="Second meta input file for cross- instead of self-similarity")
    p.add_argument("input", help="Meta file of input to process")
    p.add_argument("output", help="Image output file")
    a = p.parse_args(argv)
    sr = _rate(a, a.input)
    span = _span(a.span_start, a.span_stop, sr)
    _require(_non_empty(span), "Span is empty")
    corr = secs_to_frames(a.length, sr)
    _require(corr > 0, "Correlation duration is zero")
    if not a.no_norm and a.dir is None:
        p.print_usage(out)
        return 1
    ss.verbose = a.verbose
    b = ss.ConfigBuilder()
    b.metaInput, b.metaInput2, b.imageOutput, b.span, b.corrLen = a.input, a.input2, a.output, span, corr
    b.decimation, b.temporalWeight, b.colors = a.decim, a.temp, ss.color_scheme(a.colors)
    b.colorWarp, b.colorCeil, b.colorInv, b.normalize = a.color_warp, a.color_ceil, a.color_inv, not a.no_norm
    if a.dir is not None:
        b.databaseFolder = a.dir
    return _go(ss.SelfSimilarity, b.build(), lambda _: out.write("  Done.\n\n"), out)


def feature_cross(argv: List[str], out=sys.stdout) -> int:
    p = _parser(f"{NAME} -y")
    p.add_argument("-d", "--dir", default=None, help="Database directory (required for normalization file)")
    p.add_argument("--temp", type=float, default=0.5, help="Temporal weight (0 to 1, default 0.5)")
    p.add_argument("--span1-start", type=float, default=None, help="Correlation begin in first file (secs)")
    p.add_argument("--span1-stop", type=float, default=None, help="Correlation end in first file (secs)")
    p.add_argument("--span2-start", type=float, default=None, help="Correlation begin in second file (secs)")
    p.add_argument("--span2-stop", type=float, default=None, help="Correlation end in second file (secs)")
    p.add_argument("--boost-max", type=float, default=8.0, help="Maximum loudness boost factor (default 8)")
    p.add_argument("--no-norm", action="store_true", help="Do not apply feature normalization")
    p.add_argument("input1", help="Meta file of first input to process")
    p.add_argument("input2", help="Meta file of second input to process")
    p.add_argument("output", help="Audio output file")
    a = p.parse_args(argv)
    sr1 = _rate(a, a.input1)
    sr2 = _rate(a, a.input2)
    span1, span2 = _span(a.span1_start, a.span1_stop, sr1), _span(a.span2_start, a.span2_stop, sr2)
    _require(_non_empty(span1), "Span1 is empty")
    _require(_non_empty(span2), "Span2 is empty")
    if not a.no_norm and a.dir is None:
        p.print_usage(out)
        return 1
    b = cs.ConfigBuilder()
    b.metaInput1, b.metaInput2, b.audioOutput, b.span1, b.span2 = a.input1, a.input2, a.output, span1, span2
    b.temporalWeight, b.maxBoost, b.normalize = a.temp, a.boost_max, not a.no_norm
    if a.dir is not None:
        b.databaseFolder = a.dir
    return _go(cs.CrossSimilarity, b.build(), lambda _: out.write("  Done.\n\n"), out)


def feature_stats(argv: List[str], out=sys.stdout) -> int:
    p = _parser(f"{NAME} --stats")
    p.add_argument("-d", "--dir", required=True, help="Database directory")
    a = p.parse_args(argv)
    out.write("Starting stats... \n")
    paths = sorted(os.path.join(a.dir, n) for n in os.listdir(a.dir) if n.endswith("_feat.aif"))

    def success(spans):
        out.write("  Success.\n")
        fst.write_norms(a.dir, spans)
        out.write("Done.\n")

    return _go(fst.FeatureStats, paths, success, out)


def main(argv: Optional[List[str]] = None, out=sys.stdout) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    modes = {"-c": feature_corr, "--correlate": feature_corr, "-s": feature_segm, "--segmentation": feature_segm,
             "-x": feature_self, "--selfsimilarity": feature_self, "-y": feature_cross, "--crosssimilarity": feature_cross,
             "--stats": feature_stats}
    if not argv or argv[0] not in modes:
        if argv and argv[0] in ("-f", "--feature"):
            out.write("Feature extraction runs SuperCollider (scsynth) and is not part of the B200 engine; use the reference for it.\n")
            return 1
        out.write(f"Usage: {NAME} [-c|--correlate] [-s|--segmentation] [-x|--selfsimilarity] [-y|--crosssimilarity] [--stats] <options>\n"
                  "  -c  Find best correlation with database\n  -s  Find segmentation breaks with a file\n"
                  "  -x  Create an image of the self similarity matrix\n  -y  Create a cross-similarity vector file\n"
                  "  --stats  Statistics from feature database\n")
        return 1
    try:
        return modes[argv[0]](argv[1:], out)
    except ValueError as e:
        out.write(f"{e}\n")
        return 1
    except SystemExit as e:          # argparse: a missing / malformed option, like scopt's parse failure -> exit 1
        return 1 if e.code else 0


if __name__ == "__main__":
    sys.exit(main())
