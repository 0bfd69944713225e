"""GPU parity of the SHARDED path against the single-rank path, on the hardware the driver runs `-m gpu` on (one B200):
two processes share cuda:0 and talk over gloo (NCCL refuses two ranks on one device; the collectives' payloads are
identical).  What bench.py reports as `parity.equals_n1` on 2 / 4 / 8 GPUs is asserted here as a test:

  * K_Means.fit / fit_mix with `process_group`: identical centres on every rank (seeding included, random_state=None
    included), labels / centres equal to the single-rank fit where the initialisation is index-based;
  * the row-sharded round (labels, top-5 indices, packed records, ONE all-gather, vote) equals the single-rank round
    bit for bit; the vocabulary-column-sharded scoring (all-gather + k-way merge) equals the unsharded lists.
"""
import os
import socket
import traceback

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gather_equal(t, group, world):
    """True when the tensor is bitwise identical on every rank (NaN payloads included)."""
    from scd_b200 import dist as sdist
    bits = t.contiguous().view(torch.int32) if t.dtype == torch.float32 else t.contiguous()
    out = torch.empty((world,) + tuple(bits.shape), dtype=bits.dtype, device=bits.device)
    sdist.all_gather_into(out, bits, group)
    return all(torch.equal(out[0], out[r]) for r in range(1, world))


def _worker(rank, world, port, errq):
    try:
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        import datetime
        dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
        torch.cuda.set_device(0)
        from scd_b200 import dist as sdist, kmeans, naming, synth
        group = dist.group.WORLD
        cfg = synth.Config('mr', 9001, 24, 3000, 5)
        data = synth.make(cfg, d=256)
        X, Xc, W = data['X'], data['Xc'], data['W']
        lo, hi = sdist.shard_bounds(cfg.n, world, rank)
        Xs = X[lo:hi].cuda()

        # ---- fit, index-based initialisations: the sharded fit is the single-rank fit
        for init, seed in (('random', 7), ('first', None)):
            one = kmeans.K_Means(k=cfg.k, max_iterations=6, n_init=1, init=init, random_state=seed)
            one.fit(X.cuda())
            sh = kmeans.K_Means(k=cfg.k, max_iterations=6, n_init=1, init=init, random_state=seed, process_group=group)
            sh.fit(Xs)
            assert _gather_equal(sh.cluster_centers_, group, world), f'{init}: centres differ between the ranks'
            assert sh.n_iter_ == one.n_iter_
            assert torch.allclose(sh.cluster_centers_, one.cluster_centers_, atol=1e-5, rtol=1e-5, equal_nan=True)
            agree = (sh.labels_ == one.labels_[lo:hi]).float().mean().item()
            assert agree > 0.999, f'{init}: labels agree on {agree:.4f}'
            assert abs(float(sh.inertia_) - float(one.inertia_)) < 1e-4 * max(1.0, float(one.inertia_))

        # ---- fit, k-means++ with random_state=None: every rank must still end with the same centres
        sh = kmeans.K_Means(k=cfg.k, max_iterations=5, n_init=2, init='k-means++', random_state=None, process_group=group)
        sh.fit(Xs)
        assert _gather_equal(sh.cluster_centers_, group, world), 'k-means++: centres differ between the ranks'
        assert not torch.isnan(sh.cluster_centers_).any()
        one = kmeans.K_Means(k=cfg.k, max_iterations=5, n_init=2, init='k-means++', random_state=0)
        one.fit(X.cuda())
        assert float(sh.inertia_) < 1.25 * float(one.inertia_)            # same quality of seeding as the single-rank draw

        # ---- fit_mix: labelled rows replicated, counted once.  k == number of labelled classes: no random draw at all
        y = data['y']
        lab_mask = y < 6
        L, lt = X[lab_mask][:300], y[lab_mask][:300]
        U = X[~lab_mask]
        ulo, uhi = sdist.shard_bounds(len(U), world, rank)
        one = kmeans.K_Means(k=6, max_iterations=5, n_init=1, random_state=1)
        one.fit_mix(U.cuda(), L.cuda(), lt.cuda())
        sh = kmeans.K_Means(k=6, max_iterations=5, n_init=1, random_state=1, process_group=group)
        sh.fit_mix(U[ulo:uhi].cuda(), L.cuda(), lt.cuda())
        assert _gather_equal(sh.cluster_centers_, group, world)
        assert torch.allclose(sh.cluster_centers_, one.cluster_centers_, atol=1e-5, rtol=1e-5)
        assert sh.n_iter_ == one.n_iter_ == len(lt)
        n_l = len(lt)
        assert torch.equal(sh.labels_[:n_l], one.labels_[:n_l])
        assert (sh.labels_[n_l:] == one.labels_[n_l + ulo:n_l + uhi]).float().mean().item() > 0.999
        # and with more clusters than labelled classes the sharded k-means++ draw keeps the ranks in step
        sh = kmeans.K_Means(k=12, max_iterations=4, n_init=1, random_state=None, process_group=group)
        sh.fit_mix(U[ulo:uhi].cuda(), L.cuda(), lt.cuda())
        assert _gather_equal(sh.cluster_centers_, group, world)

        # ---- the row-sharded round equals the single-rank round bit for bit
        C0 = data['C0'].cuda()
        lab1 = torch.empty(cfg.n, dtype=torch.int64, device='cuda')
        kmeans._estep(X.cuda(), C0, lab1, None)
        vocab = naming.Vocabulary(W.cuda())
        _, idx1 = naming.score_topk(Xc, vocab, k=5)
        names1, counts1, distinct1, rows1, _ = naming.vote_device(idx1, lab1, cfg.k, 5, 20)
        lab_s = torch.empty(hi - lo, dtype=torch.int64, device='cuda')
        kmeans._estep(Xs, C0, lab_s, None)
        _, idx_s = naming.score_topk(Xc[lo:hi], vocab, k=5)
        assert torch.equal(lab_s, lab1[lo:hi]) and torch.equal(idx_s, idx1[lo:hi])
        gat = sdist.RowGather(cfg.n, (6,), torch.int32, torch.device('cuda'), group)
        names, counts, distinct, rows, ovf = sdist.sharded_vote(lab_s, idx_s, 5, cfg.k, 20, gat)
        assert int(ovf.item()) == 0
        assert torch.equal(names, names1) and torch.equal(counts, counts1) and torch.equal(distinct, distinct1) and torch.equal(rows, rows1)
        assert torch.equal(gat.full[:, 0].long(), lab1) and torch.equal(gat.full[:, 1:].long(), idx1)

        # ---- vocabulary column-sharded scoring: all-gather + k-way merge = the unsharded lists (logits and softmax)
        clo, chi = sdist.shard_bounds(cfg.v, world, rank)
        shard = naming.Vocabulary(W[:, clo:chi].cuda(), col_offset=clo)
        fb = naming._feats_bf16(Xc)
        for softmax in (False, True):
            want_v, want_i = naming.score_topk(Xc, vocab, k=5, softmax=softmax)
            got_v, got_i = sdist.sharded_score_topk(fb, shard, 5, softmax, group)
            assert torch.equal(got_i, want_i)
            assert torch.allclose(got_v, want_v, atol=1e-6 if softmax else 1e-5, rtol=1e-5)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f'rank {rank}:\n{traceback.format_exc()}')
        raise


@pytest.mark.timeout(600)
def test_two_ranks_on_one_gpu_match_the_single_rank_path():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(570)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    alive = [p for p in procs if p.is_alive()]
    for p in alive:
        p.kill()
    assert not msgs, '\n'.join(msgs)
    assert not alive, 'a rank did not finish'
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
