"""CPU tests of the decorated p-value post-processing (row f4; reference R/DecoratedPvalue.R:198-304)."""
import numpy as np

from geneticscre_b200 import decorated, synth


def test_hand_checked_case():
    """n = 6 (3 cases), sub-path carries patient 0, the added gene carries patients 1 and 4 -> one case, one control.
    Method 1, table vt[i][j] = 10 i + j: real score vt[2][1] = 21.  Redrawing 2 carriers among patients 1..5 (2 cases, 3
    controls): P(2 cases) = 1/10 -> vt[3][0] = 30, P(1) = 6/10 -> 21, P(0) = 3/10 -> vt[1][2] = 12; p = P(score >= 21) = 0.7."""
    vt = np.array([[10.0 * i + j for j in range(7)] for i in range(7)])
    pos1 = np.array([1, 0, 0, 0, 0, 0], bool)
    pos2 = np.array([0, 1, 0, 0, 1, 0], bool)
    z = np.zeros(6, bool)
    r = decorated.compute_decorated_pvalue(pos1, z, pos2, z, 3, 3, 1, vt)
    assert (r.score, r.cases1, r.controls1, r.cases2, r.controls2) == (21.0, 1, 0, 1, 1)
    assert abs(r.decorated_pvalue - 0.7) < 1e-12


def test_monte_carlo_converges_to_exact():
    nc, nt = 60, 70
    rng = np.random.default_rng(3)
    vt = synth.make_value_table(nc, nt)
    rows = rng.random((4, nc + nt)) < 0.08
    for method in (1, 2):
        exact = decorated.decorated_pvalues_for_path(rows, [1, -1, 1, -1], nc, nt, method, vt)
        mc = decorated.decorated_pvalues_for_path(rows, [1, -1, 1, -1], nc, nt, method, vt, n_permutations=40000, rng=np.random.default_rng(11))
        assert len(exact) == len(mc) == 6
        for e, m in zip(exact, mc):
            assert e["direction"] == m["direction"] and e["score"] == m["score"]
            sigma = np.sqrt(max(e["decorated_pvalue"] * (1 - e["decorated_pvalue"]), 1e-6) / 40000)
            assert abs(e["decorated_pvalue"] - m["decorated_pvalue"]) < 5 * sigma + 1e-9


def _r_style_stratified(pos1, neg1, pos2, neg2, nc, nt, method, vt, strata, n_perm, rng):
    """Literal restatement of the reference's stratified loop (R/DecoratedPvalue.R:200-293) with explicit index sampling -
    the independent check of the count-based implementation in decorated.py."""
    n = nc + nt
    i_pos1, i_neg1 = set(np.flatnonzero(pos1)), set(np.flatnonzero(neg1))
    i_pos2 = [i for i in np.flatnonzero(pos2) if i not in i_pos1]
    i_neg2 = [i for i in np.flatnonzero(neg2) if i not in i_neg1]
    case = lambda idx: sum(1 for i in idx if i < nc)
    ctrl = lambda idx: sum(1 for i in idx if i >= nc)
    cp1, cn1, tp1, tn1 = case(i_pos1), ctrl(i_neg1), ctrl(i_pos1), case(i_neg1)
    cp2, cn2, tp2, tn2 = case(i_pos2), ctrl(i_neg2), ctrl(i_pos2), case(i_neg2)

    def sc(a, b, c, d):  # (case_pos, control_pos, case_neg, control_neg) of the redrawn part
        if method == 1:
            return vt[cp1 + a + cn1 + c, tp1 + b + tn1 + d]
        return vt[cp1 + a, tp1 + b] + vt[cn1 + c, tn1 + d]

    score = sc(cp2, tp2, cn2, tn2)
    inds_1 = i_pos1 | i_neg1
    groups = list(dict.fromkeys(strata.tolist()))
    pools = {g: [i for i in range(n) if strata[i] == g and i not in inds_1] for g in groups}
    k_pos = {g: sum(1 for i in i_pos2 if i in set(pools[g])) for g in groups}
    k_neg = {g: sum(1 for i in i_neg2 if i in set(pools[g])) for g in groups}
    hits = 0
    for _ in range(n_perm):
        s_pos, s_neg = [], []
        for g in groups:
            pool = list(pools[g])
            take = list(rng.choice(pool, size=k_pos[g], replace=False)) if k_pos[g] else []
            s_pos += take
            rest = [i for i in pool if i not in set(take)]
            s_neg += list(rng.choice(rest, size=k_neg[g], replace=False)) if k_neg[g] else []
        a, c = case(s_pos), ctrl(s_neg)
        hits += sc(a, len(s_pos) - a, c, len(s_neg) - c) >= score
    return score, hits / n_perm


def test_stratified_single_stratum_equals_unstratified_for_method1():
    """Method 1 has no negative part: with one stratum the pool is "everything outside the sub-path", as without strata."""
    nc, nt = 40, 50
    rng = np.random.default_rng(5)
    vt = synth.make_value_table(nc, nt)
    rows = rng.random((3, nc + nt)) < 0.1
    a = decorated.decorated_pvalues_for_path(rows, [1, 1, 1], nc, nt, 1, vt)
    b = decorated.decorated_pvalues_for_path(rows, [1, 1, 1], nc, nt, 1, vt, strata=np.zeros(nc + nt, dtype=int))
    for x, y in zip(a, b):
        assert abs(x["decorated_pvalue"] - y["decorated_pvalue"]) < 1e-12 and x["score"] == y["score"]


def test_stratified_exact_and_monte_carlo_against_the_reference_loop():
    nc, nt = 30, 36
    n = nc + nt
    rng = np.random.default_rng(8)
    vt = synth.make_value_table(nc, nt)
    strata = rng.integers(0, 3, size=n)
    for method in (1, 2):
        pos1, pos2 = rng.random(n) < 0.10, rng.random(n) < 0.12
        neg1 = (rng.random(n) < 0.08) if method == 2 else np.zeros(n, bool)
        neg2 = (rng.random(n) < 0.10) if method == 2 else np.zeros(n, bool)
        exact = decorated.compute_decorated_pvalue(pos1, neg1, pos2, neg2, nc, nt, method, vt, strata=strata)
        mc = decorated.compute_decorated_pvalue(pos1, neg1, pos2, neg2, nc, nt, method, vt, n_permutations=60000,
                                                rng=np.random.default_rng(21), strata=strata)
        score, ref = _r_style_stratified(pos1, neg1, pos2, neg2, nc, nt, method, vt, strata, 6000, np.random.default_rng(22))
        assert exact.score == mc.score == score
        p = exact.decorated_pvalue
        assert abs(p - mc.decorated_pvalue) < 5 * np.sqrt(max(p * (1 - p), 1e-6) / 60000) + 1e-9
        assert abs(p - ref) < 5 * np.sqrt(max(p * (1 - p), 1e-6) / 6000) + 1e-9


def test_stratified_rejects_a_stratum_that_is_too_small():
    import pytest

    vt = np.zeros((5, 5))
    pos1 = np.array([0, 0, 0, 0, 1, 1, 1, 0], bool)     # stratum 1 (patients 4..7) keeps one free patient
    pos2 = np.array([0, 0, 0, 0, 0, 0, 0, 1], bool)
    neg2 = np.array([0, 0, 0, 0, 0, 0, 0, 1], bool)     # ... but two carriers (one positive, one negative) to redraw in it
    z = np.zeros(8, bool)
    with pytest.raises(ValueError):
        decorated.compute_decorated_pvalue(pos1, z, pos2, neg2, 4, 4, 2, vt, strata=np.array([0, 0, 0, 0, 1, 1, 1, 1]))
