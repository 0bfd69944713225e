"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL on
the GPUs, gloo in the CPU tests) for the two exchange steps the path really has:

  * k-means: rows block-sharded, one packed all-reduce of ``[K*D sums | K counts | inertia]`` per
    iteration (``scd_b200.kmeans.K_Means(process_group=...)``);
  * naming: vocabulary column-sharded, local fused top-k per rank (global column indices), all-gather of
    the ``[N, k]`` (value, index) lists (+ per-row max / sum-exp when the softmax is wanted) and a k-way
    merge kernel.  Row-sharded naming needs no exchange at all.
"""
from __future__ import annotations

import torch


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous block ``[lo, hi)`` of ``total`` items owned by ``rank`` (ceil-sized blocks, last one ragged)."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def merge_topk_parts(parts, k: int, softmax: bool, scale: float = 100.0):
    """parts: list of ``(vals [N,k] scaled logits, idx [N,k] int64 global, row_max [N], row_sumexp [N])`` from
    ``naming.name_topk_raw(..., softmax=False, want_stats=True)`` on each vocabulary shard -> merged
    ``(vals, idx)`` (softmax probabilities when ``softmax``)."""
    from . import _lib
    lib = _lib.load()
    vals = torch.stack([p[0] for p in parts]).contiguous()
    idx = torch.stack([p[1] for p in parts]).contiguous()
    n = int(vals.shape[1])
    out_v = torch.empty(n, k, dtype=torch.float32, device=vals.device)
    out_i = torch.empty(n, k, dtype=torch.int64, device=vals.device)
    pmax = psum = None
    if softmax:
        pmax = torch.stack([p[2] for p in parts]).contiguous()
        psum = torch.stack([p[3] for p in parts]).contiguous()
    _lib.check(lib.scd_topk_merge(vals.data_ptr(), idx.data_ptr(), _lib.ptr(pmax), _lib.ptr(psum), len(parts), n, k,
                                  float(scale), int(bool(softmax)), out_v.data_ptr(), out_i.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), 'scd_topk_merge')
    return out_v, out_i


def sharded_score_topk(feats_bf16: torch.Tensor, vocab_shard, k: int, softmax: bool, group=None, scale: float = 100.0):
    """Vocabulary-column-parallel scoring: every rank holds all rows and its own ``naming.Vocabulary`` shard
    (``col_offset`` = first global column).  Local fused top-k -> all-gather -> k-way merge; every rank
    returns the full ``(vals [N,k], idx [N,k])``."""
    import torch.distributed as dist
    from . import naming
    vals, idx, rmax, rsum = naming.name_topk_raw(feats_bf16, vocab_shard, k, False, scale, want_stats=True)
    world = dist.get_world_size(group)
    if world == 1:
        return merge_topk_parts([(vals, idx, rmax, rsum)], k, softmax, scale)
    gv = [torch.empty_like(vals) for _ in range(world)]
    gi = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(gv, vals, group=group)
    dist.all_gather(gi, idx, group=group)
    if softmax:
        gm = [torch.empty_like(rmax) for _ in range(world)]
        gs = [torch.empty_like(rsum) for _ in range(world)]
        dist.all_gather(gm, rmax, group=group)
        dist.all_gather(gs, rsum, group=group)
    else:
        gm = gs = [None] * world
    return merge_topk_parts(list(zip(gv, gi, gm, gs)), k, softmax, scale)
