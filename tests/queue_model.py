"""A second, independent restatement of the reference's match queues -- test infrastructure, like the oracle.

oracle/sgz_oracle.c restates FeatureCorrelationImpl.body() in C with arrays; this module restates the SELECTION part of the
same method (FeatureCorrelationImpl.scala:113-150 entryHasSpace / lowestSim / addMatch, :190-246 loop A, :250-393 loops B and
C, :399-400 the merge) in Python with a small ordered-set class that mimics scala.collection.immutable.SortedSet[Match]
(MatchMinOrd, FeatureCorrelation.scala:75-77), written from the Scala source and not from the C.  It takes the similarity and
boost CURVES from the oracle (so the arithmetic is the oracle's) and replays the queues on them; tests/test_queue_model.py
checks that both restatements return the same matches bit for bit over a few hundred randomly drawn configurations.  The two
documented deviations of the oracle from the JVM (nothing is read past the written part of the temp
files, a file shorter than the punch yields no offset) are the model's too."""
from __future__ import annotations

import math
import struct
from typing import List, Optional

import numpy as np

F32 = np.float32


def java_float_key(x) -> int:
    """total order of java.lang.Float.compare as an integer key: -0.0 < 0.0, NaN (canonical) above +Infinity"""
    x = F32(x)
    if np.isnan(x):
        bits = 0x7FC00000
    else:
        bits = struct.unpack("<i", struct.pack("<f", float(x)))[0]
    return bits if bits >= 0 else bits ^ 0x7FFFFFFF


class Match:
    __slots__ = ("sim", "file", "start", "stop", "boost_in", "boost_out")

    def __init__(self, sim, file, start, stop, boost_in, boost_out):
        self.sim, self.file, self.start, self.stop = F32(sim), file, int(start), int(stop)
        self.boost_in, self.boost_out = F32(boost_in), F32(boost_out)


class MinOrdSet:
    """SortedSet[Match](MatchMinOrd): iteration from the LARGEST sim down (compare(a, b) = b.sim compare a.sim); two matches
    are the same element when their sims compare equal; adding an element that is already there changes nothing."""

    def __init__(self, items: Optional[List[Match]] = None):
        self.items: List[Match] = list(items or [])      # descending Float.compare order

    def __len__(self):
        return len(self.items)

    def _find(self, sim):
        k = java_float_key(sim)
        for i, m in enumerate(self.items):
            km = java_float_key(m.sim)
            if km == k:
                return i, True
            if km < k:
                return i, False
        return len(self.items), False

    def add(self, m: Match):
        i, found = self._find(m.sim)
        if not found:
            self.items.insert(i, m)

    def remove(self, m: Match):
        i, found = self._find(m.sim)
        if found:
            del self.items[i]

    def last(self) -> Match:
        return self.items[-1]

    def take(self, n: int) -> "MinOrdSet":
        return MinOrdSet(self.items[:max(n, 0)])


def spacing(a_start, a_stop, b_start, b_stop) -> int:
    """SpanUtil.spacing (impl/SpanUtil.scala:38-43)"""
    return b_start - a_stop if a_start < b_start else a_start - b_stop


def search(O, op, files) -> List[dict]:
    """FeatureCorrelationImpl.body() with the curves taken from oracle.corr_curve (O = the oracle module, op = CorrParams)."""
    step = op.step_size

    def full_to_feat(x):
        return (x + step // 2) // step

    def feat_to_full(i):
        return i * step

    w_in = full_to_feat(op.punch_in[1]) - full_to_feat(op.punch_in[0])
    has_out = op.punch_out is not None
    w_out = full_to_feat(op.punch_out[1]) - full_to_feat(op.punch_out[0]) if has_out else 0
    min_punch, max_punch = full_to_feat(op.min_punch), full_to_feat(op.max_punch)
    num_matches, num_per_file, min_spacing = op.num_matches, op.num_per_file, op.min_spacing

    all_prio = MinOrdSet()
    state = {"entry": MinOrdSet(), "last": None}

    def entry_has_space():
        return len(state["entry"]) < min(num_matches - len(all_prio), num_per_file)

    def lowest_sim():
        if len(state["entry"]):
            return state["entry"].last().sim
        if len(all_prio):
            return all_prio.last().sim
        return F32(0.0)

    def add_match(m: Match):
        entry, last = state["entry"], state["last"]
        if last is not None and spacing(m.start, m.stop, last.start, last.stop) < min_spacing:
            if last.sim < m.sim:            # primitive comparison: false for NaN
                entry.remove(last)
                entry.add(m)
                state["last"] = m
        else:
            entry.add(m)
            if len(entry) > num_per_file:
                entry.remove(entry.last())
            state["last"] = m

    for idx, f in enumerate(files):
        state["entry"], state["last"] = MinOrdSet(), None
        n = f.shape[0]
        left = n - (min_punch if has_out else 0)
        n_a = left - w_in + 1 if left >= w_in else 0           # Q6: a file shorter than the punch yields no offset
        if n_a > 0:
            sim_in, boost_in = O.corr_curve(op, f, 0, 0)
        t_in, t_in_off, t_in_open = [], 0, False
        for off in range(n_a):
            sim, boost = F32(sim_in[off]), F32(boost_in[off])
            if has_out:
                if t_in_open or entry_has_space() or sim > lowest_sim():
                    if not t_in_open:
                        t_in_off, t_in_open = off, True
                    t_in.append((sim, boost))
            elif entry_has_space() or sim > lowest_sim():
                add_match(Match(sim, idx, feat_to_full(off), feat_to_full(off + w_in), boost, 1.0))
        if has_out and t_in_open:
            po_off0 = t_in_off + min_punch
            left = n - po_off0
            if left >= w_out:
                sim_out, boost_out = O.corr_curve(op, f, 1, po_off0)      # loop B starts its ring buffer at poOff0
                assert len(sim_out) == left - w_out + 1
                for k, (in_sim, b_in) in enumerate(t_in):
                    pi_off = t_in_off + k
                    low, hs = lowest_sim(), entry_has_space()
                    if in_sim > F32(low * low):
                        # Q5: only what loop B has written is searched
                        left2 = min(len(sim_out) - k, max_punch - min_punch + 1)
                        po_off = pi_off + min_punch
                        for j in range(max(left2, 0)):
                            out_sim, b_out = F32(sim_out[k + j]), F32(boost_out[k + j])
                            prod = float(F32(in_sim * out_sim))
                            sim = F32(math.sqrt(prod)) if prod >= 0.0 else F32(np.nan)      # math.sqrt(x).toFloat
                            if hs or sim > low:
                                add_match(Match(sim, idx, feat_to_full(pi_off), feat_to_full(po_off), b_in, b_out))
                                low, hs = lowest_sim(), entry_has_space()
                            po_off += 1
        for m in state["entry"].items:                     # allPrio ++= entryPrio; take(numMatches)
            all_prio.add(m)
        if len(all_prio) > num_matches:
            all_prio = all_prio.take(num_matches)

    return [dict(sim=float(m.sim), file=m.file, start=m.start, stop=m.stop, boostIn=float(m.boost_in),
                 boostOut=float(m.boost_out)) for m in all_prio.items]


class MaxOrdSet(MinOrdSet):
    """SortedSet[Break](BreakMaxOrd): iteration from the SMALLEST sim up (FeatureSegmentation.scala:60-62); `last` is the
    largest.  Implemented on the descending list of MinOrdSet read backwards."""

    def last(self):
        return self.items[0]

    def ascending(self):
        return self.items[::-1]


def segmentation(O, sp, file) -> List[dict]:
    """FeatureSegmentationImpl.body() (:52-142) on the oracle's similarity curve: entryHasSpace / highestSim / addBreak"""
    step = sp.step_size

    def full_to_feat(x):
        return (x + step // 2) // step

    half = full_to_feat(sp.corr_len)
    win = 2 * half
    n = file.shape[0]
    af_start = max(0, full_to_feat(sp.span_start)) if sp.span_start is not None else 0
    af_stop = min(n, full_to_feat(sp.span_stop)) if sp.span_stop is not None else n
    af_len = af_stop - af_start
    _, curve = O.segm_run(sp, file, want_curve=True)
    prio = MaxOrdSet()
    last = None
    # the first read takes min(left, winLen) frames, every later one a single frame (:103-107): a span shorter than the
    # window still yields ONE offset, evaluated on the freshly allocated (zero-filled) rest of the buffer
    n_off = af_len - win + 1 if af_len >= win else (1 if af_len > 0 else 0)
    for off in range(n_off):
        sim = F32(curve[off])
        highest = prio.last().sim if len(prio) else F32(0.0)
        if len(prio) < sp.num_breaks or sim < highest:
            b = Match(sim, 0, (af_start + off + half) * step, 0, 0, 0)      # Break(sim, pos): pos in `start`
            if last is not None and (b.start - last.start) < sp.min_spacing:
                if last.sim > b.sim:
                    prio.remove(last)
                    prio.add(b)
                    last = b
            else:
                prio.add(b)
                if len(prio) > sp.num_breaks:
                    prio.remove(prio.last())
                last = b
    return [dict(sim=float(b.sim), pos=b.start) for b in prio.ascending()]
