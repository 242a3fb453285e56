"""Decorated p-values on the device: gcre_exec_decorated_exact against
  * exact rational arithmetic (math.comb / Fraction) on small cohorts - an evaluation that shares nothing with the kernel, and
  * the host implementation (geneticscre_b200/decorated.py, exact mode) on cohort-sized inputs.
Tolerance: 1e-12 relative on the p-value (stated per the north star); counts and the f64 score are bit-exact."""
from fractions import Fraction
from math import comb

import numpy as np
import pytest

from geneticscre_b200 import decorated, synth

pytestmark = pytest.mark.gpu


def exact_pvalue(pos1, neg1, pos2, neg2, n_cases, n_ctrls, method, vt):
    n = n_cases + n_ctrls
    is_case = np.arange(n) < n_cases
    pos2 = pos2 & ~pos1
    neg2 = neg2 & ~neg1
    cp1, tp1 = int((pos1 & is_case).sum()), int((pos1 & ~is_case).sum())
    cn1, tn1 = int((neg1 & ~is_case).sum()), int((neg1 & is_case).sum())
    cp2, cn2 = int((pos2 & is_case).sum()), int((neg2 & ~is_case).sum())
    kp, kn = int(pos2.sum()), int(neg2.sum())
    gp, bp = n_cases - cp1, n_ctrls - tp1
    gn, bn = n_ctrls - cn1, n_cases - tn1

    def sc(a, b):
        if method == 1:
            return vt[cp1 + a + cn1 + b, tp1 + (kp - a) + tn1 + (kn - b)]
        return vt[cp1 + a, tp1 + (kp - a)] + vt[cn1 + b, tn1 + (kn - b)]

    score = sc(cp2, cn2)
    tot = Fraction(0)
    for a in range(max(0, kp - bp), min(kp, gp) + 1):
        pa = Fraction(comb(gp, a) * comb(bp, kp - a), comb(gp + bp, kp))
        for b in range(max(0, kn - bn), min(kn, gn) + 1):
            if sc(a, b) >= score:
                tot += pa * Fraction(comb(gn, b) * comb(bn, kn - b), comb(gn + bn, kn))
    return float(tot), float(score)


def random_splits(rng, n, count, density):
    out = []
    for _ in range(count):
        rows = rng.random((4, n)) < density
        rows[1] &= rng.random(n) < 0.5  # sparser negative parts
        rows[3] &= rng.random(n) < 0.3
        out.append(tuple(rows))
    return out


@pytest.mark.parametrize("method", [1, 2])
def test_device_matches_exact_rational_arithmetic(engine, method):
    nc, nt = 37, 44  # method 1 adds "cases" of both parts: keep the sums inside the table
    rng = np.random.default_rng(5)
    vt = synth.make_value_table(nc, nt) if method == 2 else synth.make_test_table(nc + nt, nc + nt, 3)
    ex = engine.JoinExec("method1" if method == 1 else "method2", nc, nt, 1)
    ex.setValueTable(vt)
    splits = random_splits(rng, nc + nt, 12, 0.12)
    got = decorated.decorated_exact_device(ex, splits)
    for sp, g in zip(splits, got):
        want_p, want_s = exact_pvalue(*sp, nc, nt, method, vt)
        assert g.score == want_s
        assert abs(g.decorated_pvalue - want_p) <= 1e-12 * max(want_p, 1e-300), (g.decorated_pvalue, want_p)


@pytest.mark.parametrize("method", [1, 2])
def test_device_matches_host_at_cohort_size(engine, method):
    nc, nt = 1500, 1700
    n = nc + nt
    rng = np.random.default_rng(11)
    vt = synth.make_value_table(nc, nt)
    ex = engine.JoinExec("method1" if method == 1 else "method2", nc, nt, 1)
    ex.setValueTable(vt)
    splits = random_splits(rng, n, 10, 0.03)
    splits.append((np.zeros(n, bool), np.zeros(n, bool), rng.random(n) < 0.02, np.zeros(n, bool)))  # empty sub-path
    splits.append((rng.random(n) < 0.05, np.zeros(n, bool), np.zeros(n, bool), np.zeros(n, bool)))  # gene adds nothing: p = 1
    got = decorated.decorated_exact_device(ex, splits)
    for sp, g in zip(splits, got):
        w = decorated.compute_decorated_pvalue(*sp, nc, nt, method, vt, n_permutations=None)
        assert (g.cases1, g.controls1, g.cases2, g.controls2) == (w.cases1, w.controls1, w.cases2, w.controls2)
        assert g.score == w.score
        # the host sums gammaln-based probabilities (cancellation ~ n ln n * 1e-16); the device uses the exact ratio recurrence
        assert abs(g.decorated_pvalue - w.decorated_pvalue) <= 2e-11 * max(w.decorated_pvalue, 1e-300), (g.decorated_pvalue, w.decorated_pvalue)
    assert got[-1].decorated_pvalue == pytest.approx(1.0, abs=1e-12)


def test_path_level_wrapper_uses_the_device(engine):
    nc, nt = 300, 340
    w = synth.make_workload(nc, nt, 60, 200, 4, seed=8, max_path_length=3, real_table=True, max_freq=0.1, zero_frac=0.2)
    genes = synth.unpack_bits(w.gene_bits[:4], nc + nt)
    signs = np.array([1, -1, 1, -1])
    ex = engine.JoinExec("method2", nc, nt, 1)
    ex.setValueTable(w.value_table)
    host = decorated.decorated_pvalues_for_path(genes, signs, nc, nt, 2, w.value_table)
    dev = decorated.decorated_pvalues_for_path(genes, signs, nc, nt, 2, w.value_table, device_exec=ex)
    assert len(host) == len(dev) == 6
    for h, d in zip(host, dev):
        assert (h["direction"], h["subpath1"], h["subpath2"], h["score"], h["cases1"], h["cases2"]) == (d["direction"], d["subpath1"], d["subpath2"], d["score"], d["cases1"], d["cases2"])
        assert abs(h["decorated_pvalue"] - d["decorated_pvalue"]) <= 1e-11 * max(h["decorated_pvalue"], 1e-300)
