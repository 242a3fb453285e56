"""Value-table generation (reference: getValuesTable, R/Utils.R:137-159), pinned to references the
builder did not derive from the same formulas:

  * EXACT: big-integer combinatorics (math.comb), two-sided sums as exact rationals, one log at the end - the mathematically
    exact evaluation of the reference's expression, exact ties between outcomes included;
  * scipy.stats.hypergeom.pmf with the tie guard of R's fisher.test (relErr = 1 + 1e-7).

Tolerance (stated per the north star: 1e-12 relative): |got - want| <= 1e-12 * max(1, |want|) * g(n), g(n) = max(1, n ln n / 1000).
Entries are -log p: near p = 1 they are ~0 and only an absolute error is meaningful, hence max(1, |want|); the log-factorials
every double implementation starts from carry an absolute rounding error of ~1.1e-16 * n ln n (one ulp of lgamma(n)), each
log-probability combines nine of them and the error passes 1:1 into -log p, hence g(n) (1 up to n ~ 200; 92 at n = 10,000, where
the device generator and the numpy restatement were measured 6.2e-11 apart).
"""
import math
from math import comb

import numpy as np
import pytest

from geneticscre_b200 import synth


def tol(n):
    return 1e-12 * max(1.0, n * math.log(max(n, 2)) / 1000.0)


def exact_table(nc, nt):
    n = nc + nt
    t = np.full((nc + 1, nt + 1), np.nan)
    for i in range(n + 1):
        lo, hi = max(0, i - nt), min(i, nc)
        den = comb(n, i)
        num = [comb(nc, x) * comb(nt, i - x) for x in range(lo, hi + 1)]
        for k, x in enumerate(range(lo, hi + 1)):
            s = sum(v for v in num if v <= num[k])  # exact: ties included
            t[x, i - x] = 0.0 if s == den else -(math.log(s) - math.log(den))  # math.log takes integers of any size
    assert np.isfinite(t).all()
    return t


def scipy_table(nc, nt):
    from scipy.stats import hypergeom

    n = nc + nt
    t = np.full((nc + 1, nt + 1), np.nan)
    for i in range(n + 1):
        lo, hi = max(0, i - nt), min(i, nc)
        x = np.arange(lo, hi + 1)
        p = hypergeom.pmf(x, n, nc, i)
        two = np.array([p[p <= v * (1 + 1e-7)].sum() for v in p])  # R: sapply(prob_dist, function(x) sum(prob_dist[prob_dist <= x]))
        with np.errstate(divide="ignore"):
            t[x, i - x] = -np.log(two)
    fin = np.isfinite(t)
    t[~fin] = t[fin].max() + 1.0
    return t


def assert_tables_close(got, want, n, what):
    assert got.shape == want.shape, what
    normal = want <= 700.0  # beyond: -log of sums in the denormal range (p < 1e-304), no relative accuracy on either side
    err = np.abs(got - want)[normal] / np.maximum(np.abs(want[normal]), 1.0)
    assert err.max() <= tol(n), f"{what}: max error {err.max():.3g} > {tol(n):.3g}"
    assert (got[~normal] > 700.0).all(), what


SMALL = [(2, 2), (1, 7), (57, 131), (100, 100), (150, 170)]
MID = [(700, 300), (600, 600)]


@pytest.mark.parametrize("nc,nt", SMALL)
def test_host_table_matches_exact_arithmetic(nc, nt):
    assert_tables_close(synth.make_value_table(nc, nt), exact_table(nc, nt), nc + nt, f"numpy vs exact ({nc},{nt})")


@pytest.mark.parametrize("nc,nt", MID)
def test_host_table_matches_scipy(nc, nt):
    assert_tables_close(synth.make_value_table(nc, nt), scipy_table(nc, nt), nc + nt, f"numpy vs scipy ({nc},{nt})")


def test_exact_ties_are_counted():
    """(57, 131) has outcomes with exactly equal probabilities that are not mirror images; a plain float `<=` drops some of them
    (the error is a whole probability term, ~0.2 in -log p).  The generators must agree with exact arithmetic there."""
    nc, nt = 57, 131
    ties = 0
    for i in range(nc + nt + 1):
        num = [comb(nc, x) * comb(nt, i - x) for x in range(max(0, i - nt), min(i, nc) + 1)]
        ties += len(num) - len(set(num))
    assert ties > 0
    assert_tables_close(synth.make_value_table(nc, nt), exact_table(nc, nt), nc + nt, "ties")


@pytest.mark.gpu
@pytest.mark.parametrize("nc,nt", SMALL)
def test_device_table_matches_exact_arithmetic(engine, nc, nt):
    ex = engine.JoinExec("method1", nc, nt, 1)
    ex.generateValueTable()
    assert_tables_close(ex.getValueTable(), exact_table(nc, nt), nc + nt, f"device vs exact ({nc},{nt})")


@pytest.mark.gpu
@pytest.mark.parametrize("nc,nt", MID)
def test_device_table_matches_scipy(engine, nc, nt):
    ex = engine.JoinExec("method1", nc, nt, 1)
    ex.generateValueTable()
    assert_tables_close(ex.getValueTable(), scipy_table(nc, nt), nc + nt, f"device vs scipy ({nc},{nt})")


@pytest.mark.gpu
def test_device_table_at_config3_size(engine):
    """n = 10,000 (BASELINE config 3): against the numpy restatement, itself pinned above at the sizes exact arithmetic and
    scipy finish in seconds."""
    ex = engine.JoinExec("method1", 5000, 5000, 1)
    ex.generateValueTable()
    got = ex.getValueTable()
    assert np.isfinite(got).all()
    assert_tables_close(got, synth.make_value_table(5000, 5000), 10000, "device vs numpy (5000,5000)")
