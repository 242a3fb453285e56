"""CPU tests of the multi-GPU host logic with world_size 2 over gloo: sharding by pair count, allreduce(max) of the
permutation maxima and the top-K gather/merge reproduce the unsharded join (computed here by the CPU oracle)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers
from geneticscre_b200 import dist as gdist
from geneticscre_b200 import schedule, synth


def test_shard_bounds_balance():
    count = np.array([0, 5, 1, 0, 100, 3, 3, 0, 50, 2], dtype=np.int32)
    for n in (1, 2, 3, 4, 8, 16):
        b = gdist.shard_bounds(count, n)
        assert len(b) == n and b[0][0] == 0 and b[-1][1] == count.size
        assert all(b[i][1] == b[i + 1][0] for i in range(n - 1)) and all(x <= y for x, y in b)
    assert gdist.shard_bounds(np.zeros(0, np.int32), 3) == [(0, 0)] * 3


def _worker(rank, world, port, method, out_q):
    import torch.distributed as dist

    from geneticscre_b200 import api
    from oracle import pyoracle as po

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workload(60, 70, 120, 420, 40, seed=99, max_path_length=4, real_table=True, max_freq=0.1, zero_frac=0.3)
    ex = po.OracleExec(method, w.n_cases, w.n_ctrls, w.n_perms)
    ex.top_k = 7
    ex.setValueTable(w.value_table)
    ex.setPermutedMasks(w.perm_masks)
    full, kept = schedule.replay_levels(ex, po.UidRelSet, w, 4)
    lv = w.net.levels["4"]
    b, e = gdist.shard_bounds(lv.count, world)[rank]
    # shard = rows [b, e) of the upstream set with their own index entries
    sub0 = kept["paths3"].select(np.arange(b, e, dtype=np.int32))
    uids = po.UidRelSet(4, lv.src[b:e], lv.trg[b:e], lv.count[b:e], lv.location[b:e], lv.signs[b:e])
    part = ex.join(uids, sub0, kept["paths2"], ex.createPathSet(0))
    for s in part.scores:  # row numbers are global in the engine's sharded join
        if s.src >= 0:
            s.src += b
    merged = gdist.merge_shard_result(part, 7, api.merge_topk, api.Score, dist)
    if rank == 0:
        out_q.put(([(s.score, s.src, s.trg, s.cases, s.ctrls) for s in merged.scores], merged.permuted_scores,
                   [(s.score, s.src, s.trg, s.cases, s.ctrls) for s in full["4"].scores], full["4"].permuted_scores))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("method", ["method1", "method2"])
def test_two_rank_merge_equals_full_join(oracles, engine, method):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, method, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_scores, got_perm, want_scores, want_perm = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got_scores == want_scores
    assert np.array_equal(np.asarray(got_perm).view(np.uint64), np.asarray(want_perm).view(np.uint64))
