"""Multi-GPU decomposition of one join: shard the upstream rows, merge with ONE allreduce(max) (+ a top-K gather).

Each (upstream row, partner) pair is independent; the only cross-pair state of ``JoinExec::join`` is the per-permutation
maximum and the top-K set, which the reference already merges from independent per-thread partials
(src/methods.h:25-39).  Ranks therefore own contiguous ranges of upstream rows balanced by pair count, replicate the
operands, and combine results with ``all_reduce(MAX)`` on the float32 maxima (cast to f32 happens before the reduce; the
rounding is monotone so the result equals the reference's) and an all-gather of <= K candidates per rank.

One process per GPU; ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(count, n_shards: int):
    """Contiguous upstream-row ranges ``[(begin, end)] * n_shards`` balanced by sum(count)."""
    count = np.asarray(count)
    csum = np.cumsum(count.astype(np.int64))
    total = int(csum[-1]) if csum.size else 0
    cuts = [0]
    for s in range(1, n_shards):
        cuts.append(max(cuts[-1], int(np.searchsorted(csum, total * s / n_shards))))
    cuts.append(int(count.shape[0]))
    return [(cuts[i], max(cuts[i + 1], cuts[i])) for i in range(n_shards)]


def merge_shard_result(res, top_k: int, merge_topk, score_cls, dist, device_perm=None, n_perms=None):
    """Combine this rank's shard-local ``joined_res`` with the other ranks' (in place) and return it.

    ``device_perm``: optional float32 torch tensor already holding this rank's maxima on the communicator's device
    (the GPU path exports them there without a host round trip); otherwise ``res.permuted_scores`` is used.
    """
    import torch

    world = dist.get_world_size()
    if device_perm is None:
        t = torch.from_numpy(np.asarray(res.permuted_scores, dtype=np.float32).copy())
        if dist.get_backend() == "nccl":
            t = t.cuda()
    else:
        t = device_perm
    if t.numel():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the one data-path collective of a join
    n = t.numel() if n_perms is None else n_perms
    res.permuted_scores = t[:n].to(torch.float64).cpu().numpy()
    # top-K lists as fixed-size tensors (K + 1 rows of score, src, trg, cases, ctrls; unused rows hold NaN)
    rows = torch.full((top_k + 1, 5), float("nan"), dtype=torch.float64)
    for i, sc in enumerate(res.scores[: top_k + 1]):
        rows[i] = torch.tensor([sc.score, sc.src, sc.trg, sc.cases, sc.ctrls], dtype=torch.float64)
    rows = rows.to(t.device)
    gathered = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(gathered, rows)
    lists = []
    for gt in gathered:
        gt = gt.cpu().numpy()
        lists.append([score_cls(float(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4])) for r in gt if not np.isnan(r[1])])
    res.scores = merge_topk(lists, top_k)
    return res
