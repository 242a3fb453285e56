"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import os

import numpy as np

from geneticscre_b200 import schedule, synth
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    path_length = int(z["path_length"])
    net = synth.network_from_edges(int(z["n_genes"]), z["edges_src"], z["edges_trg"], z["edges_sign"], ents2=z["ents2"],
                                   max_path_length=max(path_length, 3))
    w = synth.Workload(int(z["n_cases"]), int(z["n_ctrls"]), int(z["n_perms"]), z["gene_bits"], z["gene_bits2"], z["perm_masks"],
                       z["value_table"], net)
    want = {}
    for lvl in z["levels"]:
        lvl = str(lvl)
        sc = [po.Score(float(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4])) for r in z[f"scores_{lvl}"]]
        want[lvl] = po.JoinedRes(sc, z[f"perm_{lvl}"])
    kept = {k[5:]: z[k] for k in z.files if k.startswith("kept_")}
    return w, str(z["method"]), path_length, int(z["top_k"]), want, kept


def run_schedule(exec_cls, uid_cls, w, method, path_length, top_k, use_int_matrices=False, int_perms=False, **exec_kw):
    ex = exec_cls(method, w.n_cases, w.n_ctrls, w.n_perms, **exec_kw)
    ex.top_k = top_k
    ex.setValueTable(w.value_table)
    if int_perms:
        n = w.n_patients
        bits = np.unpackbits(w.perm_masks.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
        is_case = np.zeros(n, dtype=bool)
        is_case[: w.n_cases] = True
        ex.setPermutedCases((bits == is_case[None, :]).astype(np.int32))
    else:
        ex.setPermutedMasks(w.perm_masks)
    res, kept = schedule.replay_levels(ex, uid_cls, w, path_length, use_int_matrices=use_int_matrices)
    return res, {k: v.to_numpy() for k, v in kept.items()}, ex


def level_operands(w, kept, level):
    """numpy (paths0, paths1) of a level given the kept path sets (unpadded rows), following src/wrapper.cpp:225-276."""
    net, m_words = w.net, kept["paths1"].shape[1]
    w64 = synth.words_for(w.n_patients)

    def sel(bits, idx):
        out = np.zeros((len(idx), m_words), dtype=np.uint64)
        out[:, :w64] = bits[idx]
        return out

    if level == "1a":
        return np.zeros((net.n_genes, m_words), np.uint64), sel(w.gene_bits, net.data_idx["1a"])
    if level == "1b":
        return np.zeros((net.ents2.shape[0], m_words), np.uint64), sel(w.gene_bits2, net.data_idx["1b"])
    if level == "2":
        return kept["paths1"], sel(w.gene_bits, net.data_idx["2"])
    if level == "3":
        return kept["paths2"], sel(w.gene_bits, net.data_idx["3"])
    if level == "4":
        return kept["paths3"], kept["paths2"]
    return kept["paths3"], kept["paths3"]


def make_recompute(w, method, kept, level):
    """(src, trg) -> (score, cases, ctrls) recomputed from the raw operand rows (restates R/CheckResults.R:50-73)."""
    m = 1 if method == "method1" else 2
    w64 = synth.words_for(w.n_patients)
    p0, p1 = level_operands(w, kept, level)
    lv = w.net.levels[level]

    def rec(src, trg):
        a, b = p0[src], p1[trg]
        if m == 2:
            pl, s = lv.path_length, lv.signs
            sign = s[src] if pl > 3 else (s[trg] if pl < 3 else (-1 if s[src] + s[trg] == 0 else 1))
            if sign != 1:  # src/methods.h:140-142: downstream halves swap
                b = np.concatenate([b[w64:], b[:w64]])
        return po.row_score(m, w.n_cases, w64, a | b, w.value_table)

    return rec


def assert_same_results(got, want, what=""):
    """Exact equality incl. ties - valid when both sides follow the engine's deterministic tie rule (oracle vs CUDA)."""
    g = [(s.score, s.src, s.trg, s.cases, s.ctrls) for s in got.scores]
    e = [(s.score, s.src, s.trg, s.cases, s.ctrls) for s in want.scores]
    assert g == e, f"{what}: top-K differs\n got  {g}\n want {e}"
    gp = np.asarray(got.permuted_scores, np.float64)
    wp = np.asarray(want.permuted_scores, np.float64)
    assert gp.shape == wp.shape, what
    bad = np.nonzero(gp.view(np.uint64) != wp.view(np.uint64))[0]
    assert bad.size == 0, f"{what}: {bad.size} perm maxima differ at {bad[:8]}: {gp[bad[:8]]} vs {wp[bad[:8]]}"
