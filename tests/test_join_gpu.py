"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on identical seeded inputs.

Integer counts, kept path-set words, f64 scores and f32-rounded permutation maxima must be bit-exact.
"""
import math

import numpy as np
import pytest

import helpers
from geneticscre_b200 import _lib, synth
from test_oracle import GOLDENS, _kat, check_kat

pytestmark = pytest.mark.gpu



class _Precount(int):
    """KERNEL_SPARSE with the pre-counted-partner kernels forced on (GCRE_TEST_PRECOUNT=1); plain KERNEL_SPARSE forces the
    delta kernels, so both families run on every shape whatever the cost model would pick."""


class _NoSplit(int):
    """KERNEL_SPARSE with the split-carrier form for <= 512 permutations switched off (GCRE_TEST_NO_SPLIT=1), so the
    one-word-per-lane kernel - delta form, counts handed from level to level - runs on the small shapes too."""


class _Thresholded(_NoSplit):
    """As _NoSplit, with the thresholded look-ups (join_sparse.cuh, THR kernels) forced onto joins of any size and both methods
    (GCRE_THR=1); every other variant forces them off (GCRE_THR=0), so both look-up stages run on every shape."""


class _SharedMasks(int):
    """KERNEL_SPARSE with the split-carrier kernels' shared-memory form (join_sparse_sc.cuh, PTS: the patient-major masks staged
    in shared memory by a bulk copy, one 1,024-thread CTA per SM) forced onto every join of <= 128 permutations it fits
    (GCRE_SC_PTS=1); every other variant forces it off, so both forms run on the small shapes."""


SPARSE_PC = _Precount(_lib.KERNEL_SPARSE)
SPARSE_NOSPLIT = _NoSplit(_lib.KERNEL_SPARSE)
SPARSE_THR = _Thresholded(_lib.KERNEL_SPARSE)
SPARSE_PTS = _SharedMasks(_lib.KERNEL_SPARSE)
KERNELS = [pytest.param(_lib.KERNEL_DENSE, id="dense"), pytest.param(_lib.KERNEL_SPARSE, id="sparse"), pytest.param(SPARSE_PC, id="sparse_pc"),
           pytest.param(SPARSE_NOSPLIT, id="sparse_nosplit"), pytest.param(SPARSE_THR, id="sparse_thr"), pytest.param(SPARSE_PTS, id="sparse_pts")]
PC_MODES = [pytest.param("0", id="delta"), pytest.param("1", id="precount")]


@pytest.fixture(autouse=True)
def _precount_mode(request, monkeypatch):
    params = getattr(getattr(request.node, "callspec", None), "params", {})
    k = params.get("kernel")
    if isinstance(k, _Precount):
        monkeypatch.setenv("GCRE_TEST_PRECOUNT", "1")
    elif k == _lib.KERNEL_SPARSE:
        monkeypatch.setenv("GCRE_TEST_PRECOUNT", "0")
    monkeypatch.setenv("GCRE_SC_PTS", "1" if isinstance(k, _SharedMasks) else "0")
    if isinstance(k, _NoSplit):
        monkeypatch.setenv("GCRE_TEST_NO_SPLIT", "1")
    monkeypatch.setenv("GCRE_THR", "1" if isinstance(k, _Thresholded) else "0")
    if "pc" in params:
        monkeypatch.setenv("GCRE_TEST_PRECOUNT", params["pc"])



def run_engine(engine, w, method, path_length, top_k, kernel, **kw):
    class Exec(engine.JoinExec):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.kernel = kernel

    return helpers.run_schedule(Exec, engine.UidRelSet, w, method, path_length, top_k, **kw)


@pytest.mark.parametrize("method,sign", [("method1", 1), ("method1", -1), ("method2", 1), ("method2", -1)])
def test_known_answer(engine, method, sign):
    check_kat(_kat(engine.JoinExec, engine.UidRelSet, method, sign), method, sign)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", GOLDENS)
def test_golden_fixtures(engine, oracles, name, kernel):
    """BASELINE configs 1 and 2 (vignette data) + a synthetic 5-level case against outputs of the reference itself."""
    w, method, path_length, top_k, want, kept_want = helpers.load_golden(name)
    res, kept, _ = run_engine(engine, w, method, path_length, top_k, kernel)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k]), f"kept {k} differs"
    for lvl in want:
        oracles.compare_results(res[lvl], want[lvl], recompute=helpers.make_recompute(w, method, kept, lvl), what=f"{name} L{lvl}")


SHAPES = [
    # n_cases, n_ctrls, genes, edges, perms, top_k
    (4, 4, 30, 60, 1, 3),
    (32, 32, 60, 150, 128, 5),        # n = 64: one exact word, perms = one full tile
    (64, 64, 80, 200, 129, 7),        # n = 128: even word count, perms one over a tile
    (100, 93, 150, 500, 100, 10),     # W64 = 4 (vignette-like)
    (257, 300, 200, 800, 257, 12),    # W64 = 9 (odd -> padded to 10)
    (700, 724, 120, 420, 40, 4),      # W64 = 23, several k-chunks
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("method", ["method1", "method2"])
@pytest.mark.parametrize("shape", SHAPES)
def test_schedule_matches_oracle(engine, oracles, method, shape, kernel):
    nc, nt, g, e, perms, top_k = shape
    w = synth.make_workload(nc, nt, g, e, perms, seed=1000 + nc + perms, max_path_length=5, real_table=(nc % 2 == 0),
                            max_freq=0.12, zero_frac=0.3)
    want, kept_want, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 5, top_k)
    got, kept, _ = run_engine(engine, w, method, 5, top_k, kernel)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k]), f"kept {k} differs"
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} {shape} L{lvl}")
        assert got[lvl].info["kernel"] == kernel, "the requested kernel family must be the one that ran"
        if kernel == _lib.KERNEL_SPARSE and w.net.levels[lvl].n_pairs > 0 and lvl in ("1b", "4", "5"):
            # (joins that keep their rows emit the rows' counts for the next level instead, in the delta form)
            assert got[lvl].info["precounted"] == isinstance(kernel, _Precount), "GCRE_TEST_PRECOUNT must select the kernel form"
        if kernel == _lib.KERNEL_SPARSE and not isinstance(kernel, _Precount) and w.net.levels[lvl].n_pairs > 0:
            assert got[lvl].info["split_carrier"] == (perms <= 512 and not isinstance(kernel, _NoSplit)), "<= 512 permutations run the split-carrier form"
        if kernel == _lib.KERNEL_SPARSE and w.net.levels[lvl].n_pairs > 0 and got[lvl].info["split_carrier"]:
            assert got[lvl].info["shared_masks"] == (isinstance(kernel, _SharedMasks) and perms <= 128), "GCRE_SC_PTS must select the shared-memory form"
        if kernel == _lib.KERNEL_SPARSE and w.net.levels[lvl].n_pairs > 0 and not got[lvl].info["split_carrier"]:
            assert got[lvl].info["thresholded"] == isinstance(kernel, _Thresholded), "GCRE_THR must select the look-up stage"
            if isinstance(kernel, _Thresholded):
                # a positive maximum can only have been pushed by the exact path; never more visits than (pair, block) items
                blocks = (max(perms, 1) + 1023) // 1024
                assert got[lvl].info["exact_pairs"] <= w.net.levels[lvl].n_pairs * blocks
                assert got[lvl].info["exact_pairs"] > 0 or not (np.asarray(got[lvl].permuted_scores) > 0).any()


@pytest.mark.parametrize("pc", PC_MODES)
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_schedule_without_emitted_counts(engine, oracles, method, pc, monkeypatch):
    """By default a KEEP join hands the per-permutation counts of its rows to the next level (which then needs no carrier
    lists of its upstream operand); GCRE_TEST_EMIT=0 turns that off, so the in-kernel base walk and - with pre-counting
    forced - the KEEP form of the pre-counted kernels run on the full schedule too."""
    monkeypatch.setenv("GCRE_TEST_EMIT", "0")
    monkeypatch.setenv("GCRE_TEST_NO_SPLIT", "1")  # (257 permutations would otherwise take the split-carrier form)
    shape = SHAPES[4]
    nc, nt, g, e, perms, top_k = shape
    w = synth.make_workload(nc, nt, g, e, perms, seed=4242, max_path_length=5, real_table=True, max_freq=0.12, zero_frac=0.3)
    want, kept_want, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 5, top_k)
    got, kept, _ = run_engine(engine, w, method, 5, top_k, _lib.KERNEL_SPARSE)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k]), f"kept {k} differs"
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")


@pytest.mark.parametrize("host_pack", ["0", "1"])
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_int_matrix_inputs(engine, oracles, method, host_pack, monkeypatch):
    """The R-facing formats: IntegerMatrix data (PathSet::load) and CaseORControl (setPermutedCases).  Large matrices are
    packed to bits by host threads before the upload, small ones on the device; GCRE_TEST_HOST_PACK=1 forces the host path
    (3 threads) on this small input."""
    monkeypatch.setenv("GCRE_TEST_HOST_PACK", host_pack)
    monkeypatch.setenv("GCRE_HOST_PACK_THREADS", "3")
    w = synth.make_workload(90, 110, 100, 300, 50, seed=77, max_path_length=4, real_table=True, max_freq=0.1, zero_frac=0.2)
    want, kw, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 4, 6, use_int_matrices=True, int_perms=True)
    got, kg, _ = helpers.run_schedule(engine.JoinExec, engine.UidRelSet, w, method, 4, 6, use_int_matrices=True, int_perms=True)
    for k in kw:
        assert np.array_equal(kg[k], kw[k])
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")


def test_int_matrix_inputs_streamed_upload(engine, oracles, monkeypatch):
    """Inputs above 2 GB stream through two staging halves; GCRE_TEST_STREAMED_UPLOAD forces that path with small pieces
    (here 2,000 bytes: two to three rows per piece, many pieces, both halves reused)."""
    monkeypatch.setenv("GCRE_TEST_STREAMED_UPLOAD", "2000")
    w = synth.make_workload(90, 110, 100, 300, 50, seed=78, max_path_length=3, real_table=True, max_freq=0.1, zero_frac=0.2)
    want, kw, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, "method2", 3, 6, use_int_matrices=True, int_perms=True)
    got, kg, _ = helpers.run_schedule(engine.JoinExec, engine.UidRelSet, w, "method2", 3, 6, use_int_matrices=True, int_perms=True)
    for k in kw:
        assert np.array_equal(kg[k], kw[k])
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"streamed upload L{lvl}")


def test_two_execs_on_two_host_threads(engine, oracles):
    """Two execs driven from two host threads on their own streams (bench.py's e2e leg): uploads of one run beside the
    joins of the other through the shared block cache; both must match the oracle."""
    import threading

    w = synth.make_workload(300, 340, 200, 900, 200, seed=79, max_path_length=4, real_table=True, max_freq=0.08, zero_frac=0.2)
    want = {m: helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, m, 4, 8, use_int_matrices=True, int_perms=True)[0]
            for m in ("method1", "method2")}
    got, errs = {}, []

    def work(method):
        try:
            for _ in range(3):
                got[method] = helpers.run_schedule(engine.JoinExec, engine.UidRelSet, w, method, 4, 8, use_int_matrices=True, int_perms=True)[0]
        except BaseException as e:
            errs.append(e)

    th = [threading.Thread(target=work, args=(m,)) for m in ("method1", "method2")]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for m in want:
        for lvl in want[m]:
            helpers.assert_same_results(got[m][lvl], want[m][lvl], what=f"{m} L{lvl} (threaded)")


@pytest.mark.parametrize("method", ["method1", "method2"])
def test_precount_chosen_by_sampled_overlap(engine, oracles, method, monkeypatch):
    """Large method-1 score-only joins sample the overlap between partner and upstream rows to pick the kernel form;
    GCRE_TEST_PRECOUNT=sample applies that rule to every score-only join here.  Level 1b (empty upstream rows: nothing shared)
    must come out pre-counted, level 4 of this schedule (partner shares a gene with the upstream path) must not."""
    monkeypatch.setenv("GCRE_TEST_PRECOUNT", "sample")
    monkeypatch.setenv("GCRE_TEST_NO_SPLIT", "1")  # (the split-carrier form for <= 512 permutations never pre-counts)
    w = synth.make_workload(257, 300, 200, 800, 257, seed=4243, max_path_length=5, real_table=True, max_freq=0.12, zero_frac=0.3)
    want, kept_want, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 5, 8)
    got, kept, _ = run_engine(engine, w, method, 5, 8, _lib.KERNEL_SPARSE)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k]), f"kept {k} differs"
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")
    assert got["1b"].info["precounted"] is True
    assert got["4"].info["precounted"] is False


@pytest.mark.parametrize("method", ["method1", "method2"])
def test_device_side_inputs(engine, oracles, method, monkeypatch):
    """gcre_pathset_load_bits_device / gcre_exec_set_value_table_device (inputs that already sit in GPU memory, e.g. after
    an NCCL broadcast in a multi-GPU job) give the same results as the host-buffer calls."""
    import torch

    keep = []

    def to_device(arr, view):
        t = torch.from_numpy(np.ascontiguousarray(arr).view(view)).cuda()
        torch.cuda.synchronize()  # the exec launches on its own stream
        keep.append(t)
        return t

    def load_bits_via_device(self, bits):
        t = to_device(bits, np.int64)
        self.load_bits_device(t.data_ptr(), bits.shape[0], bits.shape[1])

    def table_via_device(self, table):
        t = to_device(table, np.float64)
        self.setValueTableDevice(t.data_ptr(), table.shape[0], table.shape[1])

    w = synth.make_workload(120, 131, 90, 320, 70, seed=88, max_path_length=4, real_table=True, max_freq=0.1, zero_frac=0.2)
    want, kw, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 4, 6)
    monkeypatch.setattr(engine.PathSet, "load_bits", load_bits_via_device)
    monkeypatch.setattr(engine.JoinExec, "setValueTable", table_via_device)
    got, kg, _ = helpers.run_schedule(engine.JoinExec, engine.UidRelSet, w, method, 4, 6)
    assert len(keep) >= 3
    for k in kw:
        assert np.array_equal(kg[k], kw[k])
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")


@pytest.mark.parametrize("rows", [3, 20])
def test_perm_rows_reused_or_truncated(engine, oracles, rows):
    """Fewer perm rows than iterations are reused cyclically, surplus rows ignored (src/join_base.cpp:89-90,116-123)."""
    nc, nt, iters = 20, 30, 8
    w = synth.make_workload(nc, nt, 40, 100, iters, seed=5, max_path_length=2, real_table=True, max_freq=0.2, zero_frac=0.1)
    perm = synth.make_perm_matrix(nc, nt, rows, seed=9)
    out = []
    for cls, ucls in ((oracles.OracleExec, oracles.UidRelSet), (engine.JoinExec, engine.UidRelSet)):
        ex = cls("method1", nc, nt, iters)
        ex.top_k = 5
        ex.setValueTable(w.value_table)
        ex.setPermutedCases(perm)
        p1 = ex.createPathSet(w.net.n_genes)
        p1.load_bits(w.gene_bits)
        lv = w.net.levels["1a"]
        out.append(ex.join(ucls(1, lv.src, lv.trg, lv.count, lv.location, lv.signs), ex.createPathSet(lv.n_uids), p1, ex.createPathSet(0)))
    helpers.assert_same_results(out[1], out[0])


def test_zero_iterations_and_small_topk(engine, oracles):
    """iterations = 0 is legal (test/issue-019.r); fewer pairs than top_k leaves the -inf sentinel in front."""
    w = synth.make_workload(10, 12, 12, 20, 0, seed=3, max_path_length=3, real_table=True, max_freq=0.3, zero_frac=0.0)
    want, _, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, "method2", 3, 1000)
    got, _, _ = helpers.run_schedule(engine.JoinExec, engine.UidRelSet, w, "method2", 3, 1000)
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=lvl)
        assert got[lvl].scores[0].src == -1 and got[lvl].scores[0].score == -math.inf
        assert got[lvl].permuted_scores.shape == (0,)


def test_prechecks_raise_like_reference(engine):
    """src/join_base.cpp:196-200 -> logic_error / out_of_range ("assertion")."""
    ex = engine.JoinExec("method1", 5, 5, 4)
    ex.setValueTable(np.zeros((6, 6)))
    ex.setPermutedMasks(np.zeros((4, 1), dtype=np.uint64))
    p0, p1 = ex.createPathSet(3), ex.createPathSet(2)
    U = engine.UidRelSet
    with pytest.raises(_lib.GcreAssertion):  # uids.size() != paths0.size
        ex.join(U(2, [0, 0], [0, 0], [1, 1], [0, 0], [1, 1]), p0, p1, ex.createPathSet(0))
    with pytest.raises(_lib.GcreOutOfRange):  # location + count - 1 >= paths1.size
        ex.join(U(2, [0] * 3, [0] * 3, [1, 2, 0], [0, 1, 0], [1, 1]), p0, p1, ex.createPathSet(0))
    with pytest.raises(_lib.GcreAssertion):  # paths_res.size not in {0, total}
        ex.join(U(2, [0] * 3, [0] * 3, [1, 1, 0], [0, 1, 0], [1, 1]), p0, p1, ex.createPathSet(5))
    with pytest.raises(_lib.GcreAssertion):  # PathSet::load row-count check (src/gcre_paths.h:60)
        p0.load(np.zeros((2, 10), dtype=np.int32))
    with pytest.raises(_lib.GcreOutOfRange):  # PathSet::select index check (src/gcre_paths.h:85)
        p0.select([0, 3])
    with pytest.raises(_lib.GcreAssertion):  # JoinExec ctor check (src/join_base.cpp:47)
        engine.JoinExec("method1", 0, 5, 4)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_sharded_join_merges_to_full(engine, oracles, method, kernel):
    """Upstream-row shards merged by max / top-K merge equal the unsharded join (the multi-GPU decomposition)."""
    w = synth.make_workload(120, 130, 150, 600, 70, seed=21, max_path_length=4, real_table=True, max_freq=0.1, zero_frac=0.3)
    want, _, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 4, 8)
    ex = engine.JoinExec(method, w.n_cases, w.n_ctrls, w.n_perms)
    ex.kernel = kernel
    ex.top_k = 8
    ex.setValueTable(w.value_table)
    ex.setPermutedMasks(w.perm_masks)
    from geneticscre_b200 import schedule

    _, kept = schedule.replay_levels(ex, engine.UidRelSet, w, 3)
    lv = w.net.levels["4"]
    uids = engine.UidRelSet(4, lv.src, lv.trg, lv.count, lv.location, lv.signs)
    cuts = [0, lv.n_uids // 3, lv.n_uids // 3, (2 * lv.n_uids) // 3, lv.n_uids]  # includes an empty shard
    parts = [ex.join(uids, kept["paths3"], kept["paths2"], ex.createPathSet(0), uid_range=(a, b)) for a, b in zip(cuts[:-1], cuts[1:]) if True]
    perm = np.max(np.stack([p.permuted_scores for p in parts]), axis=0)
    merged = engine.joined_res(engine.merge_topk([p.scores for p in parts], 8), perm)
    helpers.assert_same_results(merged, want["4"], what="sharded")
    assert sum(p.info["pairs"] for p in parts) == lv.n_pairs


def test_pathset_row_access(engine):
    ex = engine.JoinExec("method2", 40, 37, 1)
    ps = ex.createPathSet(3)
    row = np.arange(4, dtype=np.uint64) + np.uint64(1 << 12)  # (n = 77: the second word of a half holds 13 patients)
    ps.set(1, row)
    assert np.array_equal(ps[1], row)
    assert not ps[0].any()
    sel = ps.select([1, 1, 2])
    assert np.array_equal(sel.to_numpy(), np.stack([row, row, np.zeros(4, np.uint64)]))
    with pytest.raises(IndexError):
        ps[3]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("method", ["method1", "method2"])
@pytest.mark.parametrize("top_k", [9, 70])
def test_topk_launch_plan_paths(engine, oracles, method, kernel, top_k, monkeypatch):
    """Prefix launch, fixed candidate budget, budget overflow and the safe redo give the same top-K and maxima
    (top_k <= 64 on the sparse kernel: one launch with the self-tightening threshold instead of the prefix).

    GCRE_TEST_SMALL_PLAN shrinks the plan sizes so a few-thousand-pair join takes every path; the value table rises
    with the carrier count, which makes late pairs keep beating the threshold (the overflow case)."""
    monkeypatch.setenv("GCRE_TEST_SMALL_PLAN", "1")
    w = synth.make_workload(80, 85, 160, 700, 33, seed=31, max_path_length=4, real_table=False, max_freq=0.15, zero_frac=0.2)
    i, j = np.meshgrid(np.arange(w.n_cases + 1), np.arange(w.n_ctrls + 1), indexing="ij")
    w.value_table = (i * 1.0 + j * 0.5 + ((i * 7 + j * 3) % 5) * 0.01).astype(np.float64)
    want, _, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 4, top_k)
    got, _, _ = run_engine(engine, w, method, 4, top_k, kernel)
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")
    assert got["4"].info["launches"] >= 3


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_many_permutations_span_several_blocks(engine, oracles, method, kernel):
    """More than 1,024 permutations: the sparse kernel walks several 1,024-permutation blocks per unit (maxima flushed
    between blocks, true scores / kept rows only on the first), the dense kernel many 128-permutation tiles."""
    w = synth.make_workload(40, 45, 60, 170, 2100, seed=404, max_path_length=4, real_table=True, max_freq=0.15, zero_frac=0.2)
    want, kept_want, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 4, 6)
    got, kept, _ = run_engine(engine, w, method, 4, 6, kernel)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k])
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")


@pytest.mark.parametrize("perms", [300, 96])
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_dense_carrier_rows(engine, oracles, method, kernel, perms):
    """Common variants (up to half of the patients carry each gene): partner deltas of several hundred carriers exercise
    the plane-batch overflow (> 248 carriers between flushes) and the carrier-queue drain of the sparse kernel."""
    w = synth.make_workload(600, 630, 40, 110, perms, seed=505, max_path_length=4, real_table=True, max_freq=0.5, zero_frac=0.0)
    assert np.unpackbits(w.gene_bits.view(np.uint8), axis=1).sum(axis=1).max() > 400
    want, kept_want, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 4, 6)
    got, kept, _ = run_engine(engine, w, method, 4, 6, kernel)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k])
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")
        assert got[lvl].info["kernel"] == kernel


@pytest.mark.parametrize("perms", [300, 100])
@pytest.mark.parametrize("pc", PC_MODES)
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_wide_carrier_indices(engine, oracles, method, pc, perms, monkeypatch):
    """Cohorts above 65,535 patients use 32-bit carrier indices in the sparse kernel's lists; GCRE_TEST_WIDE_CARRIERS
    forces that code path on a cohort the oracle can check."""
    monkeypatch.setenv("GCRE_TEST_WIDE_CARRIERS", "1")
    # (both permutation counts take the split-carrier kernel, join_sparse_sc.cuh: 4 words for 100, 16 words for 300)
    w = synth.make_workload(300, 310, 120, 400, perms, seed=606, max_path_length=5, real_table=True, max_freq=0.3, zero_frac=0.2)
    want, kept_want, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, method, 5, 7)
    got, kept, _ = run_engine(engine, w, method, 5, 7, _lib.KERNEL_SPARSE)
    for k in kept_want:
        assert np.array_equal(kept[k], kept_want[k])
    for lvl in want:
        helpers.assert_same_results(got[lvl], want[lvl], what=f"{method} L{lvl}")
        assert got[lvl].info["kernel"] == _lib.KERNEL_SPARSE


@pytest.mark.parametrize("kernel", KERNELS)
def test_empty_and_all_zero_count_joins(engine, oracles, kernel):
    """No upstream rows at all, and upstream rows that all have count 0 with R's location -1: sentinel only, zero maxima."""
    ex = engine.JoinExec("method2", 9, 8, 5)
    ex.kernel = kernel
    ex.top_k = 3
    ex.setValueTable(synth.make_value_table(9, 8))
    ex.setPermutedMasks(synth.make_perm_masks(9, 8, 5, seed=1))
    U = engine.UidRelSet
    p1 = ex.createPathSet(4)
    for uids, p0 in ((U(4, [], [], [], [], []), ex.createPathSet(0)),
                     (U(4, [1, 2], [3, 4], [0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [1, -1]), ex.createPathSet(2))):
        r = ex.join(uids, p0, p1, ex.createPathSet(0))
        assert [(s.score, s.src, s.trg) for s in r.scores] == [(-math.inf, -1, -1)]
        assert r.permuted_scores.tolist() == [0.0] * 5 and r.info["pairs"] == 0


def test_resident_join_index_matches_one_shot(engine):
    """UidRelSet.make_resident (join index kept on the device) gives the same results as passing the arrays per join,
    also for a shard of the upstream rows."""
    w = synth.make_workload(50, 60, 90, 300, 40, seed=77, max_path_length=4, real_table=True, max_freq=0.15, zero_frac=0.2)
    ex = engine.JoinExec("method2", w.n_cases, w.n_ctrls, w.n_perms)
    ex.top_k = 5
    ex.setValueTable(w.value_table)
    ex.setPermutedMasks(w.perm_masks)
    from geneticscre_b200 import schedule

    _, kept = schedule.replay_levels(ex, engine.UidRelSet, w, 3, only=())
    lv = w.net.levels["4"]
    zero = ex.createPathSet(0)
    a = engine.UidRelSet(4, lv.src, lv.trg, lv.count, lv.location, lv.signs)
    b = engine.UidRelSet(4, lv.src, lv.trg, lv.count, lv.location, lv.signs).make_resident(ex)
    for rng in (None, (lv.n_uids // 4, lv.n_uids // 2)):
        ra = ex.join(a, kept["paths3"], kept["paths2"], zero, uid_range=rng)
        rb = ex.join(b, kept["paths3"], kept["paths2"], zero, uid_range=rng)
        helpers.assert_same_results(rb, ra, what=f"resident vs one-shot {rng}")
        assert ra.info["pairs"] == rb.info["pairs"] > 0


@pytest.mark.parametrize("perms", [300, 64])
@pytest.mark.parametrize("pc", PC_MODES)
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_sparse_queue_drain_corner_cases(engine, oracles, method, pc, perms):
    """Hand-built rows that hit the carrier-queue edge cases of the sparse kernel: exactly 64 new carriers followed by list
    entries that are all already in the upstream row (a mid-list drain, then nothing left to drain but counts still in the
    bit planes), 65 new carriers, and a partner that adds nothing."""
    n_cases, n_ctrls = 230, 170
    n = n_cases + n_ctrls
    def row(*ranges):
        v = np.zeros(n, dtype=np.int32)
        for a, b in ranges:
            v[a:b] = 1
        return v
    up = np.stack([row((200, 300)), row((10, 20), (390, 400))])
    partners = np.stack([row((0, 64), (200, 264)), row((0, 65), (200, 210)), row((200, 300)), row((150, 399))])
    table = synth.make_test_table(n_cases, n_ctrls, seed=9)
    masks = synth.make_perm_masks(n_cases, n_ctrls, perms, seed=10)
    out = []
    for cls, ucls, kern in ((oracles.OracleExec, oracles.UidRelSet, None), (engine.JoinExec, engine.UidRelSet, _lib.KERNEL_SPARSE),
                            (engine.JoinExec, engine.UidRelSet, _lib.KERNEL_DENSE)):
        ex = cls(method, n_cases, n_ctrls, perms)
        if kern is not None:
            ex.kernel = kern
        ex.top_k = 20
        ex.setValueTable(table)
        ex.setPermutedMasks(masks)
        p0, p1 = ex.createPathSet(2), ex.createPathSet(4)
        p0.load(up)
        p1.load(partners)
        res = ex.createPathSet(8)
        uids = ucls(2, [0, 1], [0, 1], [4, 4], [0, 0], [1, -1, 1, -1])
        out.append((ex.join(uids, p0, p1, res), res.to_numpy()))
    for got, rows in out[1:]:
        helpers.assert_same_results(got, out[0][0])
        assert np.array_equal(rows, out[0][1])


def test_exec_destroyed_before_its_children(engine):
    """gcre_exec_destroy orphans the exec's path sets and join indices: destroying them later is legal, using them is an error."""
    ex = engine.JoinExec("method1", 40, 41, 8)
    ps = ex.createPathSet(5)
    uid = engine.UidRelSet(1, [0], [0], [1], [0], [1]).make_resident(ex)
    ex.close()  # the wrapper's exec handle is gone; the C path set and join index are still alive
    out = np.zeros((5, 2), dtype=np.uint64)
    import ctypes as C

    assert _lib.load().gcre_pathset_download(ps._h, out.ctypes.data_as(C.POINTER(C.c_uint64))) == _lib.GCRE_ERR_ARG
    uid.release()
    del ps  # gcre_pathset_destroy on an orphan: must not touch the dead exec


def test_load_rejects_columns_beyond_n(engine):
    ex = engine.JoinExec("method1", 30, 31, 4)
    ps = ex.createPathSet(2)
    ps.load(np.ones((2, 61), dtype=np.int32))
    with pytest.raises(_lib.GcreOutOfRange):
        ps.load(np.ones((2, 62), dtype=np.int32))


@pytest.mark.parametrize("kernel", [pytest.param(_lib.KERNEL_DENSE, id="dense"), pytest.param(_lib.KERNEL_SPARSE, id="sparse")])
def test_packed_bits_beyond_n_are_dropped(engine, oracles, kernel):
    """load_bits / set with bits at patient indices >= n: cleared on entry, so both kernel families score the n real patients."""
    w = synth.make_workload(35, 35, 40, 120, 64, seed=31, max_path_length=3, real_table=True, max_freq=0.2, zero_frac=0.2)
    n = w.n_patients  # 70: the last word holds 6 patients
    dirty = w.gene_bits.copy()
    dirty[:, -1] |= np.uint64(0xFFFFFFFFFFFFFFFF) << np.uint64(n % 64)
    ex = engine.JoinExec("method2", w.n_cases, w.n_ctrls, w.n_perms)
    ex.kernel = kernel
    ps = ex.createPathSet(dirty.shape[0])
    ps.load_bits(dirty)
    got = ps.to_numpy()
    assert np.array_equal(got[:, : w.gene_bits.shape[1]], w.gene_bits) and not got[:, w.gene_bits.shape[1]:].any()
    ps.set(3, np.concatenate([dirty[5], dirty[6]]))
    assert np.array_equal(ps[3], np.concatenate([w.gene_bits[5], w.gene_bits[6]]))
    wd = synth.Workload(w.n_cases, w.n_ctrls, w.n_perms, dirty, np.ascontiguousarray(dirty[w.net.ents2]), w.perm_masks, w.value_table, w.net)
    want, _, _ = helpers.run_schedule(oracles.OracleExec, oracles.UidRelSet, w, "method2", 3, 5)
    got_r, _, _ = run_engine(engine, wd, "method2", 3, 5, kernel)
    for lvl in want:
        helpers.assert_same_results(got_r[lvl], want[lvl], what=f"dirty tail bits L{lvl}")
