"""The punch-out replay tests `cell sim > x` as `inSim * outSim >= threshold(x)` (strugatzki_b200/csrc/punchout.cuh,
cell_prod_threshold) instead of taking `math.sqrt(inSim * outSim).toFloat` (FeatureCorrelationImpl.scala:370) per cell.
This is the numpy model of that function: the same search, checked against the definition on random and edge values.
(CPU only; the kernel itself is covered by the parity tests in test_gpu_punchout.py.)"""
import numpy as np

F32 = np.float32


def sim_of_prod(p):
    with np.errstate(invalid="ignore"):
        return F32(np.sqrt(np.float64(p)))


def nudge(p, d):
    return (np.array(p, F32).view(np.int32) + np.int32(d)).view(F32)[()]


def cell_prod_threshold(a):
    a = F32(a)
    if np.isnan(a) or a == F32(np.inf):
        return F32(np.nan)
    if a < 0:
        return F32(0.0)
    with np.errstate(over="ignore"):
        pr = F32(np.float64(a) * np.float64(a))
    while True:
        if sim_of_prod(pr) > a:
            below = nudge(pr, -1)
            if not (sim_of_prod(below) > a):
                return pr
            pr = below
        else:
            pr = nudge(pr, +1)


def test_threshold_is_the_smallest_qualifying_product():
    rng = np.random.default_rng(7)
    vals = np.concatenate([
        rng.random(400, dtype=np.float32), rng.random(200, dtype=np.float32) * F32(1e-3),
        (rng.random(100, dtype=np.float32) * F32(3.0)).astype(F32),
        np.array([0.0, -0.0, 1.0, 0.5, 0.25, 1e-30, 1e-38, 1.4e-45, 3.0e38, 1.8446744e19, 2.0e19], F32)])
    for a in vals:
        t = cell_prod_threshold(a)
        assert sim_of_prod(t) > a
        assert not (sim_of_prod(nudge(t, -1)) > a)           # one Float below does not qualify any more


def test_threshold_test_equals_the_sqrt_test():
    rng = np.random.default_rng(11)
    a = rng.random(300, dtype=np.float32)
    for x in a:
        t = cell_prod_threshold(x)
        base = F32(np.float64(x) * np.float64(x))
        around = [nudge(base, d) for d in range(-6, 7) if np.array(base, F32).view(np.int32) + d >= 0]
        prods = np.array(around + list(rng.random(20, dtype=np.float32)) + [F32(0.0), F32(-0.0), F32(-1.0), F32(np.nan),
                                                                             F32(np.inf)], F32)
        with np.errstate(invalid="ignore"):
            want = np.array([sim_of_prod(p) > x for p in prods])
            got = prods >= t
        assert np.array_equal(want, got), (x, t)


def test_threshold_edge_values():
    assert np.isnan(cell_prod_threshold(np.nan)) and np.isnan(cell_prod_threshold(np.inf))    # nothing qualifies
    assert cell_prod_threshold(-0.5) == 0.0 and cell_prod_threshold(-np.inf) == 0.0            # every non-NaN sim does
    for p in (F32(0.0), F32(-0.0), F32(1e-45), F32(2.0)):
        assert (sim_of_prod(p) > F32(-0.5)) == (p >= cell_prod_threshold(-0.5))
    with np.errstate(invalid="ignore"):
        assert not (sim_of_prod(F32(-1.0)) > F32(-0.5)) and not (F32(-1.0) >= cell_prod_threshold(-0.5))
    assert cell_prod_threshold(0.0) == nudge(F32(0.0), 1) == cell_prod_threshold(-0.0)         # the smallest denormal
    big = cell_prod_threshold(F32(3.0e38))                                                     # a * a overflows
    assert big == F32(np.inf) and sim_of_prod(big) > F32(3.0e38)
