"""CPU-only, world_size 2 over gloo: the host-side logic of the sharded path - shard bounds, the packed
[K*D sums | K counts | inertia] all-reduce of the k-means M-step, and the all-gather plumbing of the
vocabulary-sharded top-k (the merge itself is a CUDA kernel, covered by the GPU tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scd_b200 import dist as sdist


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 100, 127000, 21000):
        for world in (1, 2, 3, 8):
            cuts = [sdist.shard_bounds(total, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c and a <= b and c <= d


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, k, d):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        n = 101
        X = torch.randn(n, d, generator=g)
        labels = torch.randint(0, k, (n,), generator=g)
        lo, hi = sdist.shard_bounds(n, world, rank)
        # what the M-step kernels (scd_mstep_sums + scd_pack_counts_inertia) leave on this rank's row shard
        s_sums, s_counts, s_inertia = sdist.packed_slices(k, d)
        packed = torch.zeros(k * d + k + 1)
        packed[s_sums].view(k, d).index_add_(0, labels[lo:hi], X[lo:hi])
        packed[s_counts] = torch.bincount(labels[lo:hi], minlength=k).float()
        packed[s_inertia] = float(rank + 1)
        sdist.allreduce_packed(packed, dist.group.WORLD)
        want = torch.zeros(k, d).index_add_(0, labels, X)
        assert torch.allclose(packed[s_sums].view(k, d), want, atol=1e-5)
        assert torch.equal(packed[s_counts].long(), torch.bincount(labels, minlength=k))
        assert packed[s_inertia].item() == sum(range(1, world + 1))
        # all-gather plumbing used by sharded_score_topk: every rank ends with every shard's list, in rank order
        mine = torch.full((4, 5), float(rank))
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        assert [float(t[0, 0]) for t in got] == [float(r) for r in range(world)]
        # row-sharded results (labels, top-k indices) gathered into the full [N, ...] tensor, ragged last shard
        for tail, dtype in (((), torch.int64), ((5,), torch.int64)):
            rg = sdist.RowGather(n, tail, dtype, 'cpu', dist.group.WORLD)
            assert (rg.lo, rg.hi) == (lo, hi)
            ref = torch.arange(n * max(1, int(torch.tensor(tail).prod()) if tail else 1)).view((n,) + tail)
            rg.local.copy_(ref[lo:hi])
            assert torch.equal(rg.gather(), ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize('world', [2, 3])           # 3: ragged shards (101 rows -> 34 / 34 / 33), as on 4 or 8 GPUs
def test_packed_allreduce_and_row_gather_gloo(world):
    mp.spawn(_worker, args=(world, _free_port(), 6, 16), nprocs=world, join=True)


def test_peer_exchange_is_off_without_nccl():
    """The NVLink peer-memory exchange is only chosen for NCCL groups on CUDA devices: on the CPU (and for the gloo groups
    of these tests) K_Means falls back to the packed all-reduce and `peer.available` says so."""
    from scd_b200 import kmeans, peer
    assert peer.available(None) is False
    km = kmeans.K_Means(k=3, process_group=None)
    assert km._peer_exchange(3, 8, 'cpu') is None
