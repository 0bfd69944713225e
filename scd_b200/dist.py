"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL on
the GPUs, gloo in the CPU tests) for the two exchange steps the path really has:

  * k-means: rows block-sharded, one packed all-reduce of ``[K*D sums | K counts | inertia]`` per
    iteration (``scd_b200.kmeans.K_Means(process_group=...)``);
  * naming: vocabulary column-sharded, local fused top-k per rank (global column indices), all-gather of
    the ``[N, k]`` (value, index) lists (+ per-row max / sum-exp when the softmax is wanted) and a k-way
    merge kernel.  Row-sharded naming needs no exchange at all;
  * vote: each rank packs its rows' ``[label, name_0 .. name_(k-1)]`` into int32 records, ONE all-gather
    (``RowGather``, or the peer-memory stores of ``peer.PeerExchange``) replicates them and every rank runs the
    exact vote on the gathered records (``sharded_vote``).  Per-cluster name histograms are sparse and unbounded, so "histograms add" is done
    by gathering the 24-byte rows, not K x V dense tables.
"""
from __future__ import annotations

import torch


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous block ``[lo, hi)`` of ``total`` items owned by ``rank`` (ceil-sized blocks, last one ragged)."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def packed_slices(k: int, d: int):
    """Layout of the packed fp32 M-step buffer ``[K*D sums | K counts | inertia]``."""
    return slice(0, k * d), slice(k * d, k * d + k), slice(k * d + k, k * d + k + 1)


def allreduce_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """The one exchange step of row-sharded k-means: sum the packed buffer over the ranks, in place."""
    import torch.distributed as dist
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def all_gather_into(out: torch.Tensor, inp: torch.Tensor, group=None):
    """``dist.all_gather_into_tensor`` (NCCL on the GPUs).  gloo has no all-gather for CUDA tensors - the multi-rank
    tests that share one GPU stage through the host there."""
    import torch.distributed as dist
    if dist.get_backend(group) == 'gloo':
        # gloo chunks the output along dim 0 and wants every chunk shaped like the input: exchange flat buffers
        host = torch.empty(out.numel(), dtype=out.dtype)
        dist.all_gather_into_tensor(host, inp.detach().cpu().contiguous().view(-1), group=group)
        out.copy_(host.view(out.shape))
        return out
    dist.all_gather_into_tensor(out, inp, group=group)
    return out


class RowGather:
    """All-gather of row-sharded per-row results (labels ``[n]``, top-k indices ``[n, k]``) into the full
    ``[N, ...]`` tensor on every rank.  Shards are the ceil-sized blocks of ``shard_bounds``; the ragged last
    block is padded to the common block size so a single ``all_gather_into_tensor`` does it."""

    def __init__(self, n_total: int, tail_shape, dtype, device, group=None, fill=-1):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_total = int(n_total)
        self.per = (self.n_total + self.world - 1) // self.world
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
        tail = tuple(tail_shape)
        self.full_padded = torch.full((self.world * self.per,) + tail, fill, dtype=dtype, device=device)
        self.local_padded = torch.full((self.per,) + tail, fill, dtype=dtype, device=device)
        self.local = self.local_padded[:self.hi - self.lo]      # kernels write this rank's results here
        self.full = self.full_padded[:self.n_total]

    def gather(self):
        import torch.distributed as dist
        all_gather_into(self.full_padded, self.local_padded, self.group)
        return self.full


def merge_topk_stacked(vals, idx, pmax, psum, k: int, softmax: bool, scale: float = 100.0):
    """vals / idx ``[parts, N, k]`` (scaled logits, global int64 indices) and, for the softmax, per-part
    ``row_max / row_sumexp [parts, N]`` -> merged ``(vals [N,k], idx [N,k])`` (probabilities when ``softmax``)."""
    from . import _lib
    lib = _lib.load()
    parts, n = int(vals.shape[0]), int(vals.shape[1])
    out_v = torch.empty(n, k, dtype=torch.float32, device=vals.device)
    out_i = torch.empty(n, k, dtype=torch.int64, device=vals.device)
    _lib.check(lib.scd_topk_merge(vals.data_ptr(), idx.data_ptr(), _lib.ptr(pmax) if softmax else None,
                                  _lib.ptr(psum) if softmax else None, parts, n, k, float(scale), int(bool(softmax)),
                                  out_v.data_ptr(), out_i.data_ptr(), torch.cuda.current_stream().cuda_stream), 'scd_topk_merge')
    return out_v, out_i


def merge_topk_parts(parts, k: int, softmax: bool, scale: float = 100.0):
    """parts: list of ``(vals [N,k] scaled logits, idx [N,k] int64 global, row_max [N], row_sumexp [N])`` from
    ``naming.name_topk_raw(..., softmax=False, want_stats=True)`` on each vocabulary shard -> merged
    ``(vals, idx)`` (softmax probabilities when ``softmax``)."""
    vals = torch.stack([p[0] for p in parts]).contiguous()
    idx = torch.stack([p[1] for p in parts]).contiguous()
    pmax = psum = None
    if softmax:
        pmax = torch.stack([p[2] for p in parts]).contiguous()
        psum = torch.stack([p[3] for p in parts]).contiguous()
    return merge_topk_stacked(vals, idx, pmax, psum, k, softmax, scale)


def sharded_score_topk(feats_bf16: torch.Tensor, vocab_shard, k: int, softmax: bool, group=None, scale: float = 100.0,
                       plan=None):
    """Vocabulary-column-parallel scoring: every rank holds all rows and its own ``naming.Vocabulary`` shard
    (``col_offset`` = first global column).  Local fused top-k -> all-gather -> k-way merge; every rank
    returns the full ``(vals [N,k], idx [N,k])``."""
    import torch.distributed as dist
    from . import naming
    vals, idx, rmax, rsum = naming.name_topk_raw(feats_bf16, vocab_shard, k, False, scale, want_stats=True, plan=plan)
    world = dist.get_world_size(group)
    if world == 1:
        return merge_topk_stacked(vals.unsqueeze(0), idx.unsqueeze(0), rmax.unsqueeze(0), rsum.unsqueeze(0), k, softmax, scale)
    gv = torch.empty((world,) + tuple(vals.shape), dtype=vals.dtype, device=vals.device)
    gi = torch.empty((world,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
    all_gather_into(gv, vals, group)
    all_gather_into(gi, idx, group)
    gm = gs = None
    if softmax:
        gm = torch.empty((world,) + tuple(rmax.shape), dtype=rmax.dtype, device=rmax.device)
        gs = torch.empty((world,) + tuple(rsum.shape), dtype=rsum.dtype, device=rsum.device)
        all_gather_into(gm, rmax, group)
        all_gather_into(gs, rsum, group)
    return merge_topk_stacked(gv, gi, gm, gs, k, softmax, scale)


def grid_2d(world: int, rank: int, vocab_ways: int):
    """Rank -> (row group, vocabulary group) of a ``(world // vocab_ways) x vocab_ways`` process grid: ranks that
    share a row block are consecutive, so the all-gather of their partial top-k lists stays inside the group."""
    if vocab_ways < 1 or world % vocab_ways:
        raise ValueError(f'vocab_ways={vocab_ways} does not divide the world size {world}')
    return rank // vocab_ways, rank % vocab_ways


def sharded_vote(labels_local: torch.Tensor, idx_local: torch.Tensor, top_k: int, n_clusters: int, num_common: int,
                 gather: 'RowGather', plan=None, known_name_idx=None, presorted=None):
    """Row-sharded vote: pack this rank's records into ``gather.local`` (``RowGather(n_total, (1 + top_k,), int32)``),
    all-gather once, vote on the gathered records.  Every rank returns the full result of ``naming.vote_device``.
    ``gather`` may also be a ``peer.PeerExchange``: the records are then stored straight into every rank's gathered
    array over NVLink peer memory (no NCCL launch)."""
    from . import naming
    if presorted is not None and hasattr(gather, 'gather_sorted_records'):
        # the rank's rows are already sorted by label (kmeans._MStep.sums_counts has just run on `labels_local`): exchange the
        # sorted runs + offsets and vote by walking them - no sort of the gathered records
        records, seg = gather.gather_sorted_records(presorted, idx_local, top_k)
        return naming.vote_segments(records, seg, n_clusters, num_common, known_name_idx=known_name_idx, plan=plan, n_total=gather.n_total)
    if hasattr(gather, 'gather_records'):          # peer.PeerExchange: the pack kernel's stores are the all-gather
        records = gather.gather_records(labels_local, idx_local, top_k)
    else:
        naming.pack_vote_records(labels_local, idx_local, top_k, out=gather.local)
        records = gather.gather()
    return naming.vote_records(records, n_clusters, num_common, known_name_idx=known_name_idx, plan=plan)
