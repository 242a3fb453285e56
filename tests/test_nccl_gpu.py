"""The multi-GPU path on hardware: 2 ranks, one process per GPU, NCCL.  Both decompositions of bench.py / DESIGN.md section 5:

  * upstream rows of the last level split by pair count; maxima merged with ONE all_reduce(MAX) on the device, top-K lists
    gathered and merged;
  * permutation blocks: each rank scores all pairs against its own block, each drops its maxima into its slice of a zeroed
    N x I vector and ONE all_reduce(MAX) assembles it.

Each is compared bit for bit with the single-GPU join of the whole job and with the CPU oracle.  Skips below 2 GPUs.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, method, q):
    import ctypes

    import torch
    import torch.distributed as dist

    from geneticscre_b200 import _lib, api, schedule, synth
    from geneticscre_b200 import dist as gdist
    from oracle import pyoracle as po

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lib = _lib.load()
    top_k, n_perms = 7, 300
    w = synth.make_workload(300, 340, 160, 600, n_perms, seed=77, max_path_length=4, real_table=True, max_freq=0.1, zero_frac=0.3)
    lv = w.net.levels["4"]
    uids = api.UidRelSet(4, lv.src, lv.trg, lv.count, lv.location, lv.signs)

    def prepared(masks):
        ex = api.JoinExec(method, w.n_cases, w.n_ctrls, masks.shape[0], device=rank)
        ex.set_stream(torch.cuda.current_stream().cuda_stream)
        ex.top_k = top_k
        ex.setValueTable(w.value_table)
        ex.setPermutedMasks(masks)
        _, kept = schedule.replay_levels(ex, api.UidRelSet, w, 3, only=())
        return ex, kept

    key = lambda r: ([(s.score, s.src, s.trg, s.cases, s.ctrls) for s in r.scores], np.asarray(r.permuted_scores, np.float64).view(np.uint64).tolist())

    # ---- the whole job on this GPU alone, and on the CPU oracle ----
    ex, kept = prepared(w.perm_masks)
    zero = ex.createPathSet(0)
    full = ex.join(uids, kept["paths3"], kept["paths2"], zero)
    oex = po.OracleExec(method, w.n_cases, w.n_ctrls, w.n_perms)
    oex.top_k = top_k
    oex.setValueTable(w.value_table)
    oex.setPermutedMasks(w.perm_masks)
    want, _ = schedule.replay_levels(oex, po.UidRelSet, w, 4, only=("4",))
    ok_oracle = key(full) == key(want["4"])

    # ---- upstream rows sharded: one all_reduce(MAX) of the device-resident f32 maxima + top-K gather ----
    shard = gdist.shard_bounds(lv.count, world)[rank]
    part = ex.join(uids, kept["paths3"], kept["paths2"], zero, uid_range=shard, skip_host_perm=True)
    perm_t = torch.zeros(ex.iterations, dtype=torch.float32, device="cuda")
    _lib.check(lib.gcre_exec_export_perm_max(ex._h, ctypes.c_void_p(perm_t.data_ptr()), ex.iterations))
    merged = gdist.merge_shard_result(part, top_k, api.merge_topk, api.Score, dist, device_perm=perm_t, n_perms=w.n_perms)
    ok_rows = key(merged) == key(full)

    # ---- permutation blocks: rank r scores block r; N x I vector assembled by one all_reduce(MAX) ----
    blk = n_perms // world
    mine = w.perm_masks[rank * blk:(rank + 1) * blk]
    exb, keptb = prepared(mine)
    rb = exb.join(uids, keptb["paths3"], keptb["paths2"], exb.createPathSet(0))
    allv = torch.zeros(world * exb.iterations, dtype=torch.float32, device="cuda")
    _lib.check(lib.gcre_exec_export_perm_max(exb._h, ctypes.c_void_p(allv.data_ptr() + 4 * exb.iterations * rank), exb.iterations))
    dist.all_reduce(allv, op=dist.ReduceOp.MAX)
    got = np.concatenate([allv[r * exb.iterations: r * exb.iterations + blk].cpu().numpy() for r in range(world)]).astype(np.float64)
    ok_perms = np.array_equal(got.view(np.uint64), np.asarray(full.permuted_scores[: blk * world], np.float64).view(np.uint64)) and key(rb)[0] == key(full)[0]

    flags = torch.tensor([int(ok_oracle), int(ok_rows), int(ok_perms)], dtype=torch.int32, device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(flags.cpu().tolist())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("method", ["method1", "method2"])
def test_two_ranks_over_nccl(engine, oracles, method):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, method, q)) for r in range(2)]
    for p in procs:
        p.start()
    flags = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert flags == [1, 1, 1], f"[single GPU == oracle, row shards == single GPU, perm blocks == single GPU] = {flags}"
