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
