"""CPU: the Mann-Whitney oracle (oracle/wmu_oracle.c) against the reference's own sources compiled
unmodified (oracle/_ref/libgficf_ref_wmu.so), and the restated normal cdf against scipy.

The reference's tests pin nothing for this path (tests/testthat.R:4 is commented out); GSL's cdf is
a third-party function that is absent here -> PARITY UNPINNED ON THE CDF: the restatement
(oracle/gauss_cdf.c, W. J. Cody's rational approximations as GSL implements them) is only checked to
a few ulp against an independent implementation."""
import numpy as np
import pytest
from scipy import stats
from scipy.special import ndtr

import glob
import os

from oracle.binding import WmuOracle, WmuReference

WMU_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "wmu_*.npz")))


def sc_matrix(rng, genes, cells, density=0.15, integer=False):
    """Sparse non-negative expression-like matrix: mostly zeros (one giant tie group per gene)."""
    m = rng.gamma(2.0, 50.0, size=(genes, cells)) * (rng.random((genes, cells)) < density)
    if integer:
        m = np.floor(m / 40.0)  # small integers: many tie groups
    return np.asfortranarray(m)


def test_cdf_restatement_against_scipy():
    orc = WmuOracle()
    xs = np.concatenate([np.linspace(-37.5, 8.5, 20001), [0.0, 1e-17, -1e-17, 0.66291, -0.66291, 5.6568, -5.6569]])
    p = np.array([orc.lib.gsl_cdf_ugaussian_P(float(x)) for x in xs])
    q = np.array([orc.lib.gsl_cdf_ugaussian_Q(float(x)) for x in xs])
    assert np.all(np.abs(p - ndtr(xs)) <= 1e-12 * ndtr(xs) + 1e-300)
    assert np.all(np.abs(q - ndtr(-xs)) <= 1e-12 * ndtr(-xs) + 1e-300)
    central = np.abs(xs) < 5
    assert np.max(np.abs(p[central] - ndtr(xs[central])) / ndtr(xs[central])) < 1e-14  # a few ulp where exp() is benign


@pytest.mark.parametrize("genes,n1,n2,integer", [(40, 30, 200, False), (25, 1, 50, False), (30, 64, 64, True),
                                                  (10, 500, 3000, False), (12, 7, 5, True)])
def test_oracle_equals_compiled_reference(genes, n1, n2, integer):
    if not WmuReference.available():
        pytest.skip("oracle/_ref/libgficf_ref_wmu.so not built (needs /root/reference)")
    rng = np.random.default_rng(genes + n1)
    m = sc_matrix(rng, genes, n1 + n2, integer=integer)
    m[0, :] = 3.0           # all values equal: a single tie group -> p = 1 (the reference skips the test)
    m[1, :] = 0.0
    m[2, :n1] = 0.0         # complete separation
    m[2, n1:] = rng.random(n2) + 1.0
    if genes > 4:
        m[3, :] = -m[3, :]  # negative values
        m[4, ::2] = -0.0    # -0.0 ties with +0.0
    x, y = np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:])
    want = WmuReference().wmu(x, y, nthreads=2)
    got = WmuOracle().wmu(x, y, nthreads=3)
    assert np.array_equal(got, want, equal_nan=True)  # bit for bit (same cdf restatement on both sides; log2 of a negative ratio is NaN)
    assert got[0, 0] == 1.0 and got[1, 0] == 1.0


def test_oracle_against_scipy_mannwhitneyu():
    """Independent cross-check of the statistics (tolerance: the two implementations share nothing)."""
    rng = np.random.default_rng(7)
    genes, n1, n2 = 30, 40, 160
    m = sc_matrix(rng, genes, n1 + n2, density=0.5)
    x, y = np.asfortranarray(m[:, :n1]), np.asfortranarray(m[:, n1:])
    got = WmuOracle().wmu(x, y)
    for g in range(genes):
        ref = stats.mannwhitneyu(x[g], y[g], alternative="two-sided", method="asymptotic", use_continuity=True)
        assert abs(got[g, 0] - ref.pvalue) <= 1e-9 * max(ref.pvalue, 1e-300) + 1e-15
        assert abs(got[g, 1] - np.log2((x[g] + 1).mean() / (y[g] + 1).mean())) < 1e-12


@pytest.mark.parametrize("path", WMU_GOLDEN, ids=[os.path.basename(p)[:-4] for p in WMU_GOLDEN])
def test_oracle_equals_golden_vectors(path):
    """Outputs of the reference's own sources (tests/golden/make_golden.py), committed so that the pin
    also holds where /root/reference does not exist."""
    g = np.load(path)
    assert np.array_equal(WmuOracle().wmu(g["x"], g["y"]), g["out"], equal_nan=True)
