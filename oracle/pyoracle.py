"""TEST INFRASTRUCTURE ONLY - ctypes drivers for the two CPU oracles.

* :class:`RefExec`    - the UNMODIFIED reference classes (JoinExec/PathSet/UidRelSet of /root/reference/src) compiled by
                        oracle/build_ref.sh into oracle/_ref/libgcre_ref_<isa>.so and reached through oracle/ref_shim.cpp.
* :class:`OracleExec` - our scalar C restatement (oracle/gcre_oracle.c -> oracle/libgcre_oracle.so).

Both expose the same duck-typed surface as the product's ``geneticscre_b200.api.JoinExec`` (which mirrors
src/gcre.h:103-180) so that ``geneticscre_b200.schedule.replay_levels`` can drive any of the three.
Nothing in the product imports this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


@dataclass
class Score:  # src/gcre_types.h:32-43
    score: float
    src: int
    trg: int
    cases: int
    ctrls: int


@dataclass
class JoinedRes:  # src/gcre_types.h:45-48
    scores: list
    permuted_scores: np.ndarray


class _ScoreC(C.Structure):
    _fields_ = [("score", C.c_double), ("src", C.c_int), ("trg", C.c_int), ("cases", C.c_int), ("ctrls", C.c_int)]


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_variant() -> str:
    """Best reference build this host can execute (results are identical across variants)."""
    fl = _cpu_flags()
    if {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd", "avx512_vpopcntdq", "avx512_bitalg"} <= fl:
        return "avx512"
    if {"avx2", "bmi2", "fma"} <= fl:
        return "avx2"
    return "sse42"


def ref_available(variant: str | None = None) -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libgcre_ref_{variant or ref_variant()}.so"))


def build(verbose: bool = False) -> None:
    """Build both oracles (the reference one only where /root/reference exists)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-s", "-C", HERE, "libgcre_oracle.so"], stdout=out)
    subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=out, stderr=out)


_ptr = lambda a, t: a.ctypes.data_as(C.POINTER(t))

_libc = C.CDLL(None)


class _quiet_c_stdout:
    """The reference printf()s progress lines to C stdout on every join (src/join_base.cpp:166-181); keep them out of
    the caller's stdout (bench.py must print exactly one JSON line)."""

    def __enter__(self):
        import sys

        sys.stdout.flush()
        _libc.fflush(None)
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        _libc.fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def _as(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class UidRelSet:  # src/gcre.h:49-90
    def __init__(self, path_length, src, trg, count, location, signs):
        self.path_length = int(path_length)
        self.src, self.trg, self.count = _as(src, np.int32), _as(trg, np.int32), _as(count, np.int32)
        self.location, self.signs = _as(location, np.uint32), _as(signs, np.int32)

    def size(self):
        return self.count.shape[0]

    def count_total_paths(self):
        return int(self.count.sum(dtype=np.int64))


# ---------------------------------------------------------------------------------------------------------------
# the unmodified reference
# ---------------------------------------------------------------------------------------------------------------
class _RefLib:
    _cache = {}

    @classmethod
    def get(cls, variant=None):
        variant = variant or ref_variant()
        if variant not in cls._cache:
            path = os.path.join(HERE, "_ref", f"libgcre_ref_{variant}.so")
            lib = C.CDLL(path)
            lib.ref_last_error.restype = C.c_char_p
            lib.ref_simd_label.restype = C.c_char_p
            for fn in ("ref_exec_create", "ref_pathset_create", "ref_pathset_select"):
                getattr(lib, fn).restype = C.c_void_p
            lib.ref_exec_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            lib.ref_exec_destroy.argtypes = [C.c_void_p]
            lib.ref_exec_width_ul.argtypes = [C.c_void_p]
            lib.ref_exec_iterations.argtypes = [C.c_void_p]
            lib.ref_exec_set_threads.argtypes = [C.c_void_p, C.c_int]
            lib.ref_exec_set_top_k.argtypes = [C.c_void_p, C.c_int]
            lib.ref_exec_set_value_table.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int]
            lib.ref_exec_set_perms.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]
            lib.ref_pathset_create.argtypes = [C.c_void_p, C.c_uint]
            lib.ref_pathset_destroy.argtypes = [C.c_void_p]
            lib.ref_pathset_size.argtypes = [C.c_void_p]
            lib.ref_pathset_size.restype = C.c_uint
            lib.ref_pathset_vlen.argtypes = [C.c_void_p]
            lib.ref_pathset_width_ul.argtypes = [C.c_void_p]
            lib.ref_pathset_load.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]
            lib.ref_pathset_select.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
            lib.ref_pathset_copy_out.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
            lib.ref_pathset_set_row.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_uint64)]
            lib.ref_join.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_uint), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(_ScoreC), C.c_int, C.POINTER(C.c_double)]
            cls._cache[variant] = lib
        return cls._cache[variant]


class RefPathSet:
    def __init__(self, ex, handle):
        self.ex, self.h = ex, handle
        if not handle:
            raise RuntimeError("reference: " + ex.lib.ref_last_error().decode())
        self.size = ex.lib.ref_pathset_size(handle)
        self.vlen = ex.lib.ref_pathset_vlen(handle)
        self.width_ul = ex.lib.ref_pathset_width_ul(handle)

    def __del__(self):
        if getattr(self, "h", None):
            self.ex.lib.ref_pathset_destroy(self.h)
            self.h = None

    def load(self, data):
        d = _as(data, np.int32)
        if self.ex.lib.ref_pathset_load(self.h, _ptr(d, C.c_int), d.shape[0], d.shape[1] if d.ndim == 2 else 0) != 0:
            raise RuntimeError("reference load: " + self.ex.lib.ref_last_error().decode())

    def load_bits(self, bits):
        """Write packed rows (uint64[size][W64]) into the pos half through PathSet::set."""
        bits = _as(bits, np.uint64)
        row = np.zeros(self.vlen, dtype=np.uint64)
        for r in range(self.size):
            row[:] = 0
            row[: bits.shape[1]] = bits[r]
            self.ex.lib.ref_pathset_set_row(self.h, r, _ptr(row, C.c_uint64))

    def select(self, idx):
        i = _as(idx, np.int32)
        return RefPathSet(self.ex, self.ex.lib.ref_pathset_select(self.h, _ptr(i, C.c_int), i.shape[0]))

    def to_numpy(self):
        """uint64[size][W64*method] with the reference's SIMD padding stripped."""
        out = np.zeros((self.size, max(self.vlen, 1)), dtype=np.uint64)
        if self.size:
            self.ex.lib.ref_pathset_copy_out(self.h, _ptr(out, C.c_uint64))
        w64, m = self.ex.w64, self.ex.m
        halves = [out[:, h * self.width_ul : h * self.width_ul + w64] for h in range(m)]
        return np.ascontiguousarray(np.concatenate(halves, axis=1))


class RefExec:
    """The reference's JoinExec (src/gcre.h:103-180, src/join_base.cpp)."""

    def __init__(self, method, num_cases, num_ctrls, iters, variant=None):
        self.lib = _RefLib.get(variant)
        self.method, self.num_cases, self.num_ctrls, self.iters = method, num_cases, num_ctrls, iters
        self.m = 1 if method == "method1" else 2
        self.w64 = (num_cases + num_ctrls + 63) // 64
        self._top_k, self._nthreads = 12, 0
        self.h = self.lib.ref_exec_create(method.encode(), num_cases, num_ctrls, iters, self._top_k, self._nthreads)
        if not self.h:
            raise RuntimeError("reference: " + self.lib.ref_last_error().decode())
        self.width_ul = self.lib.ref_exec_width_ul(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_exec_destroy(self.h)
            self.h = None

    top_k = property(lambda s: s._top_k, lambda s, v: (setattr(s, "_top_k", int(v)), s.lib.ref_exec_set_top_k(s.h, int(v)))[0])
    nthreads = property(lambda s: s._nthreads, lambda s, v: (setattr(s, "_nthreads", int(v)), s.lib.ref_exec_set_threads(s.h, int(v)))[0])

    def setValueTable(self, table):
        t = _as(table, np.float64)
        self.lib.ref_exec_set_value_table(self.h, _ptr(t, C.c_double), t.shape[0], t.shape[1])

    def setPermutedCases(self, perms):
        p = _as(perms, np.int32)
        with _quiet_c_stdout():
            rc = self.lib.ref_exec_set_perms(self.h, _ptr(p, C.c_int), p.shape[0], p.shape[1] if p.ndim == 2 else 0)
        if rc != 0:
            raise RuntimeError("reference perms: " + self.lib.ref_last_error().decode())

    def setPermutedMasks(self, masks):
        """The reference only accepts the int matrix: rebuild CaseORControl (1 = kept) from packed case masks."""
        n = self.num_cases + self.num_ctrls
        bits = np.unpackbits(_as(masks, np.uint64).view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
        is_case = np.zeros(n, dtype=bool)
        is_case[: self.num_cases] = True
        self.setPermutedCases((bits == is_case[None, :]).astype(np.int32))

    def createPathSet(self, size):
        return RefPathSet(self, self.lib.ref_pathset_create(self.h, int(size)))

    def join(self, uids, paths0, paths1, paths_res):
        cap = self._top_k + 1
        sc = (_ScoreC * cap)()
        perm = np.zeros(max(self.iters, 1), dtype=np.float64)
        with _quiet_c_stdout():
            n = self.lib.ref_join(self.h, uids.path_length, uids.size(), _ptr(uids.src, C.c_int), _ptr(uids.trg, C.c_int),
                                  _ptr(uids.count, C.c_int), _ptr(uids.location, C.c_uint), _ptr(uids.signs, C.c_int),
                                  uids.signs.shape[0], paths0.h, paths1.h, paths_res.h, sc, cap, _ptr(perm, C.c_double))
        if n < 0:
            raise RuntimeError("reference join: " + self.lib.ref_last_error().decode())
        return JoinedRes([Score(s.score, s.src, s.trg, s.cases, s.ctrls) for s in sc[:n]], perm[: self.iters].copy())


# ---------------------------------------------------------------------------------------------------------------
# our C restatement
# ---------------------------------------------------------------------------------------------------------------
def _oracle_lib():
    if not hasattr(_oracle_lib, "lib"):
        path = os.path.join(HERE, "libgcre_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", HERE, "libgcre_oracle.so"])
        lib = C.CDLL(path)
        lib.oracle_join.restype = C.c_int
        lib.oracle_join.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                    C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.c_int64, C.POINTER(C.c_uint64), C.c_int64,
                                    C.POINTER(C.c_uint64), C.c_int, C.POINTER(_ScoreC), C.POINTER(C.c_float)]
        lib.oracle_perm_masks.argtypes = [C.POINTER(C.c_int32), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        lib.oracle_pack_rows_i32.argtypes = [C.POINTER(C.c_int32), C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        lib.oracle_row_score.restype = C.c_double
        lib.oracle_row_score.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _oracle_lib.lib = lib
    return _oracle_lib.lib


class OraclePathSet:
    def __init__(self, ex, size):
        self.ex, self.size = ex, int(size)
        self.vlen = ex.w64 * ex.m
        self.rows = np.zeros((self.size, self.vlen), dtype=np.uint64)

    def load(self, data):
        d = _as(data, np.int32)
        if d.shape[0] != self.size:
            raise RuntimeError("assertion")
        _oracle_lib().oracle_pack_rows_i32(_ptr(d, C.c_int32), d.shape[0], d.shape[1], self.ex.w64, self.vlen, _ptr(self.rows, C.c_uint64))

    def load_bits(self, bits):
        self.rows[:] = 0
        self.rows[:, : self.ex.w64] = bits

    def select(self, idx):
        out = OraclePathSet(self.ex, len(idx))
        out.rows[:] = self.rows[np.asarray(idx, dtype=np.int64)]
        return out

    def to_numpy(self):
        return self.rows.copy()


class OracleExec:
    """CPU restatement with the same surface (oracle/gcre_oracle.c)."""

    def __init__(self, method, num_cases, num_ctrls, iters):
        self.method, self.num_cases, self.num_ctrls, self.iters = method, num_cases, num_ctrls, iters
        self.m = 1 if method == "method1" else 2
        self.w64 = (num_cases + num_ctrls + 63) // 64
        self.top_k, self.nthreads = 12, 0
        self.masks = np.zeros((iters, self.w64), dtype=np.uint64)
        self.table = np.zeros((1, 1))

    def setValueTable(self, table):
        self.table = _as(table, np.float64)

    def setPermutedCases(self, perms):
        p = _as(perms, np.int32)
        rc = _oracle_lib().oracle_perm_masks(_ptr(p, C.c_int32), p.shape[0], p.shape[1] if p.ndim == 2 else 0, self.num_cases,
                                             self.w64, self.iters, _ptr(self.masks, C.c_uint64))
        if rc != 0:
            raise RuntimeError("oracle perms")

    def setPermutedMasks(self, masks):
        self.masks = _as(masks, np.uint64)[: self.iters].copy()

    def createPathSet(self, size):
        return OraclePathSet(self, size)

    def join(self, uids, paths0, paths1, paths_res):
        total = uids.count_total_paths()
        if paths_res.size not in (0, total):
            raise RuntimeError("assertion")
        sc = (_ScoreC * (self.top_k + 1))()
        perm = np.zeros(max(self.iters, 1), dtype=np.float32)
        res_ptr = _ptr(paths_res.rows, C.c_uint64) if paths_res.size else None
        n = _oracle_lib().oracle_join(self.m, self.num_cases, self.num_ctrls, self.w64, self.iters, _ptr(self.masks, C.c_uint64),
                                      _ptr(self.table, C.c_double), self.table.shape[0], self.table.shape[1], uids.path_length,
                                      uids.size(), _ptr(uids.count, C.c_int32), _ptr(uids.location, C.c_uint32),
                                      _ptr(uids.signs, C.c_int32), _ptr(paths0.rows, C.c_uint64), paths0.size,
                                      _ptr(paths1.rows, C.c_uint64), paths1.size, res_ptr, self.top_k, sc, _ptr(perm, C.c_float))
        if n < 0:
            raise RuntimeError("assertion")
        return JoinedRes([Score(s.score, s.src, s.trg, s.cases, s.ctrls) for s in sc[:n]], perm[: self.iters].astype(np.float64))


def row_score(method_m, num_cases, w64, row, table):
    """Score of one joined row from scratch (restates R/CheckResults.R:50-73) -> (score, cases, ctrls)."""
    row = _as(row, np.uint64)
    t = _as(table, np.float64)
    a, b = C.c_int(0), C.c_int(0)
    s = _oracle_lib().oracle_row_score(method_m, num_cases, w64, _ptr(row, C.c_uint64), _ptr(t, C.c_double), t.shape[0], t.shape[1],
                                       C.byref(a), C.byref(b))
    return s, a.value, b.value


# ---------------------------------------------------------------------------------------------------------------
# parity rule
# ---------------------------------------------------------------------------------------------------------------
def compare_results(got: JoinedRes, want: JoinedRes, recompute=None, what=""):
    """Raise AssertionError unless `got` matches `want` under the A.7 rule.

    perm maxima: bit-exact (float32 values widened to double).  Top-K: identical multiset of scores (bit-exact);
    entries strictly above the K-th score must match (src,trg,cases,ctrls) as a set; entries tied with the K-th
    score must be genuine pairs with exactly that score - `recompute(src, trg) -> (score, cases, ctrls)` verifies them.
    """
    gp, wp = np.asarray(got.permuted_scores, dtype=np.float64), np.asarray(want.permuted_scores, dtype=np.float64)
    assert gp.shape == wp.shape, f"{what}: perm shape {gp.shape} vs {wp.shape}"
    bad = np.nonzero(gp.view(np.uint64) != wp.view(np.uint64))[0]
    # +0.0 and -0.0 cannot occur (maxima start at +0.0), so bit equality is the right test
    assert bad.size == 0, f"{what}: {bad.size} perm maxima differ, first r={bad[:5]} got={gp[bad[:5]]} want={wp[bad[:5]]}"
    gs, ws = got.scores, want.scores
    assert len(gs) == len(ws), f"{what}: top-K length {len(gs)} vs {len(ws)}"
    g_sorted = sorted(s.score for s in gs)
    w_sorted = sorted(s.score for s in ws)
    assert np.array_equal(np.array(g_sorted).view(np.uint64), np.array(w_sorted).view(np.uint64)), f"{what}: score multisets differ\n{g_sorted}\n{w_sorted}"
    assert all(gs[i].score <= gs[i + 1].score for i in range(len(gs) - 1)), f"{what}: scores not ascending"
    if not gs:
        return
    kth = min(s.score for s in gs if s.src >= 0) if any(s.src >= 0 for s in gs) else -math.inf
    key = lambda s: (s.score, s.src, s.trg, s.cases, s.ctrls)
    g_above = sorted(key(s) for s in gs if s.score > kth or s.src < 0)
    w_above = sorted(key(s) for s in ws if s.score > kth or s.src < 0)
    assert g_above == w_above, f"{what}: entries above the K-th score differ\n{g_above}\n{w_above}"
    if recompute is not None:
        seen = set()
        for s in gs:
            if s.src < 0:
                continue
            assert (s.src, s.trg) not in seen, f"{what}: duplicate pair {(s.src, s.trg)}"
            seen.add((s.src, s.trg))
            sc, ca, ct = recompute(s.src, s.trg)
            assert (sc, ca, ct) == (s.score, s.cases, s.ctrls), f"{what}: pair {(s.src, s.trg)} reported {key(s)} recomputed {(sc, ca, ct)}"
