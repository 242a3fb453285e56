"""GPU tests at BASELINE config-3 cohort size (10,000 patients, 1,000 permutations, W64 = 157) through size-independent
properties, plus a direct oracle comparison on a slice the scalar oracle finishes in seconds.

  * the dense (AND+POPC) and the sparse (carrier-list) kernels are independent formulations: their results must be
    bit-identical on the whole level-4 join;
  * upstream-row shards merged by max / top-K merge == the unsharded join;
  * permutation blocks: scoring [A ; B] == scoring A and B separately and concatenating (the multi-GPU weak-scaling mode);
  * a join repeated gives identical results (no order dependence in the atomics);
  * the first upstream rows against the CPU oracle, bit-exact.
"""
import numpy as np
import pytest

import helpers
from geneticscre_b200 import _lib, schedule, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    return synth.make_workload(5000, 5000, 15000, 20000, 1000, seed=20261021, max_path_length=4, real_table=True)


def _prepare(engine, w, method, kernel, masks=None, top_k=10):
    ex = engine.JoinExec(method, w.n_cases, w.n_ctrls, (masks if masks is not None else w.perm_masks).shape[0])
    ex.kernel = kernel
    ex.top_k = top_k
    ex.setValueTable(w.value_table)
    ex.setPermutedMasks(masks if masks is not None else w.perm_masks)
    _, kept = schedule.replay_levels(ex, engine.UidRelSet, w, 3, only=())
    lv = w.net.levels["4"]
    uids = engine.UidRelSet(4, lv.src, lv.trg, lv.count, lv.location, lv.signs)
    return ex, kept, uids


def _key(res):
    return [(s.score, s.src, s.trg, s.cases, s.ctrls) for s in res.scores], res.permuted_scores.view(np.uint64).tolist()


@pytest.mark.parametrize("method", ["method1", "method2"])
def test_dense_and_sparse_kernels_agree_at_full_size(engine, big, method):
    out = {}
    for kernel in (_lib.KERNEL_DENSE, _lib.KERNEL_SPARSE):
        ex, kept, uids = _prepare(engine, big, method, kernel)
        r = ex.join(uids, kept["paths3"], kept["paths2"], ex.createPathSet(0))
        assert r.info["kernel"] == kernel and r.info["pairs"] == big.net.levels["4"].n_pairs
        out[kernel] = (_key(r), kept["paths3"].to_numpy())
        r2 = ex.join(uids, kept["paths3"], kept["paths2"], ex.createPathSet(0))  # idempotence / determinism
        assert _key(r2) == out[kernel][0]
    assert out[_lib.KERNEL_DENSE][0] == out[_lib.KERNEL_SPARSE][0]
    assert np.array_equal(out[_lib.KERNEL_DENSE][1], out[_lib.KERNEL_SPARSE][1])  # kept 3-gene paths, both kernels
    assert max(out[_lib.KERNEL_SPARSE][0][1]) > 0


def test_row_shards_and_perm_blocks_compose(engine, big):
    method = "method2"
    ex, kept, uids = _prepare(engine, big, method, _lib.KERNEL_AUTO)
    zero = ex.createPathSet(0)
    full = ex.join(uids, kept["paths3"], kept["paths2"], zero)
    # upstream-row shards
    lv = big.net.levels["4"]
    from geneticscre_b200 import dist as gdist

    parts = [ex.join(uids, kept["paths3"], kept["paths2"], zero, uid_range=b) for b in gdist.shard_bounds(lv.count, 5)]
    merged_perm = np.max(np.stack([p.permuted_scores for p in parts]), axis=0)
    merged = engine.joined_res(engine.merge_topk([p.scores for p in parts], 10), merged_perm)
    assert _key(merged) == _key(full)
    # permutation blocks
    halves = []
    for blk in (slice(0, 400), slice(400, 1000)):
        exb, keptb, uidsb = _prepare(engine, big, method, _lib.KERNEL_AUTO, masks=big.perm_masks[blk])
        halves.append(exb.join(uidsb, keptb["paths3"], keptb["paths2"], exb.createPathSet(0)))
    assert np.array_equal(np.concatenate([h.permuted_scores for h in halves]).view(np.uint64), full.permuted_scores.view(np.uint64))
    assert _key(halves[0])[0] == _key(full)[0] == _key(halves[1])[0]  # the top-K does not depend on the permutations


@pytest.mark.parametrize("pts", ["0", "1"])
@pytest.mark.parametrize("method", ["method1", "method2"])
def test_few_permutations_are_a_prefix_of_many(engine, big, method, pts, monkeypatch):
    """A permutation's maximum depends on its own mask only: the first 100 / 256 permutations scored alone - split-carrier
    kernels (join_sparse_sc.cuh), counts handed from level 3 to level 4 in their layout, masks in shared memory or not - must
    give the first 100 / 256 maxima of the 1,000-permutation run (1,024-permutation kernels), the same top-K and kept rows."""
    monkeypatch.setenv("GCRE_SC_PTS", pts)
    ex, kept, uids = _prepare(engine, big, method, _lib.KERNEL_SPARSE)
    full = ex.join(uids, kept["paths3"], kept["paths2"], ex.createPathSet(0))
    assert not full.info["split_carrier"]
    p3 = kept["paths3"].to_numpy()
    for n_few in (100, 256):
        exf, keptf, uidsf = _prepare(engine, big, method, _lib.KERNEL_SPARSE, masks=big.perm_masks[:n_few])
        few = exf.join(uidsf, keptf["paths3"], keptf["paths2"], exf.createPathSet(0))
        assert few.info["split_carrier"] and few.info["shared_masks"] == (pts == "1" and n_few <= 128)
        assert np.array_equal(few.permuted_scores.view(np.uint64), full.permuted_scores[:n_few].view(np.uint64))
        assert _key(few)[0] == _key(full)[0]
        assert np.array_equal(keptf["paths3"].to_numpy(), p3)
        exf.close()


@pytest.mark.parametrize("method", ["method1", "method2"])
def test_first_rows_against_oracle_at_full_width(engine, oracles, big, method):
    ex, kept, _ = _prepare(engine, big, method, _lib.KERNEL_AUTO, top_k=12)
    lv = big.net.levels["4"]
    x = 300
    p3 = kept["paths3"].to_numpy()[:x]
    p2 = kept["paths2"].to_numpy()
    sub = ex.createPathSet(x)
    for r in range(x):
        sub.set(r, p3[r])
    got = ex.join(engine.UidRelSet(4, lv.src[:x], lv.trg[:x], lv.count[:x], lv.location[:x], lv.signs[:x]), sub, kept["paths2"], ex.createPathSet(0))
    oex = oracles.OracleExec(method, big.n_cases, big.n_ctrls, big.n_perms)
    oex.top_k = 12
    oex.setValueTable(big.value_table)
    oex.setPermutedMasks(big.perm_masks)
    o0, o1 = oex.createPathSet(x), oex.createPathSet(p2.shape[0])
    o0.rows[:] = p3
    o1.rows[:] = p2
    want = oex.join(oracles.UidRelSet(4, lv.src[:x], lv.trg[:x], lv.count[:x], lv.location[:x], lv.signs[:x]), o0, o1, oex.createPathSet(0))
    helpers.assert_same_results(got, want, what=f"{method} first {x} rows at n=10,000")
