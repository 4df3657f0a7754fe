"""The rule behind the exact re-evaluation of a punch-in search (strugatzki_b200/csrc/corr_refine.cuh), as a numpy model.

The kernel only sees APPROXIMATE file maxima (error <= err) and has to pick a threshold below which no offset can reach the
result.  The reference keeps the numMatches largest DISTINCT sims (equal sims are one TreeSet entry,
FeatureCorrelationImpl.scala:399-400), so the true bound is the K-th largest distinct exact file maximum; the kernel counts
maxima closer than tieTol = 2 err as one (their equality is only known after the re-evaluation) and subtracts margin >= 2 err.
This test checks the claim the kernel relies on: whatever the perturbation, every offset whose EXACT sim reaches the true
bound has an approximate sim above the kernel's threshold -- with repeated files (exact ties) and near-ties in the data."""
import numpy as np
import pytest


def kernel_threshold(approx_max, k, margin, tie_tol):
    """k_refine_threshold: K rounds of 'largest key below the previous one minus tieTol'"""
    keys = np.sort(np.asarray(approx_max, np.float64))[::-1]
    bound, last = np.inf, None
    for _ in range(max(k, 1)):
        below = keys[keys < bound]
        if below.size == 0:
            return -np.inf
        last = below[0]
        bound = last - tie_tol
    return last - margin


@pytest.mark.parametrize("seed", range(40))
def test_threshold_never_cuts_a_decisive_offset(seed):
    rng = np.random.default_rng(seed)
    n_files = int(rng.integers(1, 400))
    k = int(rng.choice([1, 2, 5, 20, 100]))
    err = float(rng.choice([1e-6, 1e-5, 2.5e-4]))
    margin, tie_tol = 2.0 * err * rng.uniform(1.0, 2.5), 2.0 * err
    exact = rng.uniform(0.2, 0.9, n_files)
    # repeated files (exact ties), near-ties inside and just outside the tolerance, a dense cluster
    for _ in range(int(rng.integers(0, n_files // 2 + 1))):
        a, b = rng.integers(0, n_files, 2)
        exact[a] = exact[b] + float(rng.choice([0.0, 0.0, 0.5 * err, 1.9 * err, 2.1 * err, 5 * err]))
    approx = exact + rng.uniform(-err, err, n_files)
    thr = kernel_threshold(approx, k, margin, tie_tol)
    distinct = np.unique(exact)[::-1]
    if distinct.size < k:
        assert thr == -np.inf or thr <= approx.min()          # everything is re-evaluated
        return
    true_bound = distinct[k - 1]                                # the final lowest sim is at least this
    decisive = exact >= true_bound
    assert np.all(approx[decisive] >= thr), (k, err, thr, true_bound, approx[decisive].min())
    # numPerFile = 1: within a file only offsets within the margin of the file's own approximate maximum are collected;
    # the offset that holds the exact maximum is one of them
    sims_exact = exact[:, None] - np.abs(rng.normal(0, 3 * err, (n_files, 6)))
    sims_exact[:, 0] = exact
    sims_approx = sims_exact + rng.uniform(-err, err, sims_exact.shape)
    own = sims_approx.max(axis=1)
    assert np.all(sims_approx[:, 0] >= own - margin)
