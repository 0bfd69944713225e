"""GPU parity: the size-constrained k-means drop-in (``scd_b200.sskm_constrained.K_Means``)
against the oracle restatement of ``local_utils/sskm_constrained.py`` and the fixtures the real reference module
produced (``tests/golden/kmeans_constrained.npz``; solver = stand-in, see ``oracle/constrained_oracle.py``).

Gate: size bounds hold, optimal integer total cost equals the stand-in's on the same costs, labels bit-exact where
the optimum is unique (continuous data), centroids / inertia within 1e-4 (fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import constrained_oracle as co, kmeans_oracle
from scd_b200 import kmeans, sskm_constrained as sk, synth

pytestmark = pytest.mark.gpu
ATOL = 1e-4


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, 'kmeans_constrained.npz'))


def test_int_costs_and_assignment_match_reference_fixture(golden_dir):
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['g_bounds'])
    X, C = torch.from_numpy(g['g_X']).cuda(), torch.from_numpy(g['g_C0']).cuda()
    cost = kmeans.constrained_int_costs(X, C).cpu().numpy()
    want = g['g_costs'][:240 * 6].reshape(240, 6)
    assert cost.dtype == np.int32 and np.abs(cost - want).max() <= 1          # rounding of a 1e-7 distance difference
    km = sk.K_Means(k=6, size_min=lo, size_max=hi)
    labels = torch.empty(240, dtype=torch.int64, device='cuda')
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    km._assign(X, C, labels, acc)
    lab = labels.cpu().numpy()
    sizes = np.bincount(lab, minlength=6)
    assert sizes.min() >= lo and sizes.max() <= hi and km.n_flow_solves_ == 1
    assert int(want[np.arange(240), lab].sum()) <= int(g['g_total_cost']) + 2   # optimal on the reference's own int costs
    if np.array_equal(cost, want):
        assert np.array_equal(lab, g['g_labels'])      # continuous data: unique optimum
    assert abs(acc.item() - float(g['g_inertia'])) < 1e-3


def test_fit_matches_reference_fixture(golden_dir):
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['g_bounds'])
    km = sk.K_Means(k=6, tolerance=1e-4, max_iterations=5, size_min=lo, size_max=hi, init='random', n_init=2, random_state=4,
                    n_jobs=None, pairwise_batch_size=64)
    km.fit(torch.from_numpy(g['g_X']))
    assert km.labels_.dtype == torch.int32 and km.labels_.device.type == 'cpu'
    assert np.array_equal(km.labels_.numpy(), g['fit_labels'])
    assert np.abs(km.cluster_centers_.numpy() - g['fit_centers']).max() < ATOL
    assert abs(float(km.inertia_) - float(g['fit_inertia'])) < 1e-3 and km.n_iter_ == int(g['fit_n_iter'])
    sizes = np.bincount(km.labels_.numpy(), minlength=6)
    assert sizes.min() >= lo and sizes.max() <= hi


def test_fit_mix_from_identical_seeds_matches_oracle(golden_dir):
    """k-means++ picks depend on the last bits of an fp32 cumsum, so the semi-supervised fit is compared
    from identical initial centres: one constrained restart of the oracle vs the product, same bounds."""
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['mix_bounds'])
    U, L, T = torch.from_numpy(g['mix_u']), torch.from_numpy(g['mix_l']), torch.from_numpy(g['mix_t'])
    seeds = co.K_Means(k=6).kpp(U, torch.stack([L[T == c].mean(0) for c in torch.unique(T)]), k=6, random_state=9)

    class OracleFixedSeeds(co.K_Means):
        def kpp(self, X, pre_centers=None, k=10, random_state=None):
            return seeds.clone()

    class ProductFixedSeeds(sk.K_Means):
        def kpp(self, X, pre_centers=None, k=10, random_state=None):
            return seeds.clone().cuda()

    args = dict(k=6, tolerance=1e-4, max_iterations=5, size_min=lo, size_max=hi, init='k-means++', n_init=1, random_state=9)
    ko, kp = OracleFixedSeeds(**args), ProductFixedSeeds(**args)
    ko.fit_mix(U, L, T)
    kp.fit_mix(U.cuda(), L.cuda(), T.cuda())
    assert kp.labels_.dtype == torch.int64 and kp.labels_.is_cuda
    assert torch.equal(kp.labels_.cpu(), ko.labels_)
    assert (kp.cluster_centers_.cpu() - ko.cluster_centers_).abs().max() < ATOL
    assert abs(kp.inertia_.item() - ko.inertia_.item()) < 1e-3
    assert kp.n_iter_ == ko.n_iter_ == len(T)
    u_sizes = np.bincount(kp.labels_.cpu().numpy()[len(T):], minlength=6)
    assert u_sizes.min() >= lo and u_sizes.max() <= hi and kp.n_flow_solves_ > 0


def test_loose_bounds_take_the_fast_path_and_equal_plain_kmeans():
    g = torch.Generator().manual_seed(3)
    X = synth.unit_rows(torch.randn(3000, 64, generator=g) + 3 * synth.unit_rows(torch.randn(10, 64, generator=g))[torch.randint(0, 10, (3000,), generator=g)])
    a = sk.K_Means(k=10, max_iterations=4, size_min=1, size_max=3000, init='first', n_init=1)
    b = kmeans.K_Means(k=10, max_iterations=4, init='first', n_init=1)
    a.fit(X)
    b.fit(X)
    assert a.n_flow_solves_ == 0
    dist = kmeans_oracle.pairwise_distance(X, b.cluster_centers_, None)        # labels may differ only on near-ties
    two = dist.topk(2, dim=1, largest=False).values
    ok = (two[:, 1] - two[:, 0]) > 1e-5
    assert torch.equal(a.labels_.long()[ok], b.labels_[ok])
    assert (a.cluster_centers_ - b.cluster_centers_).abs().max() < ATOL


def test_infeasible_bounds_raise_the_reference_exception():
    X = synth.unit_rows(torch.randn(50, 8, generator=torch.Generator().manual_seed(0)))
    with pytest.raises(Exception, match='min cost flow input'):
        sk.K_Means(k=4, max_iterations=1, size_min=20, size_max=30, init='first', n_init=1).fit(X)      # 4 * 20 > 50
    with pytest.raises(Exception, match='min cost flow input'):
        sk.K_Means(k=4, max_iterations=1, size_min=0, size_max=10, init='first', n_init=1).fit(X)       # 4 * 10 < 50


def test_c3_scale_round_keeps_bounds_and_total_cost():
    """C3 (20 000 x 768, K = 120) with bounds that bite: one constrained iteration; the labelling is feasible and
    its integer total cost equals the host solver's optimum on the kernel's own cost matrix."""
    cfg = synth.CONFIGS['C3']
    data = synth.make(cfg, v=256)
    X, C0 = data['X'].cuda(), data['C0'].cuda()
    lo, hi = 150, 185
    km = sk.K_Means(k=cfg.k, size_min=lo, size_max=hi)
    labels = torch.empty(cfg.n, dtype=torch.int64, device='cuda')
    acc = torch.zeros(1, dtype=torch.float64, device='cuda')
    km._assign(X, C0, labels, acc)
    lab = labels.cpu().numpy()
    sizes = np.bincount(lab, minlength=cfg.k)
    assert sizes.min() >= lo and sizes.max() <= hi and km.n_flow_solves_ == 1
    cost = kmeans.constrained_int_costs(X, C0).cpu().numpy()
    _, total, aug = sk.labels_constrained(cost, lo, hi)
    assert aug > 0 and int(cost[np.arange(cfg.n), lab].sum()) == total
    d_lab = ((data['X'] - data['C0'][labels.cpu()]) ** 2).sum(1).double().sum().item()
    assert abs(acc.item() - d_lab) < 1e-3 * max(1.0, d_lab)
