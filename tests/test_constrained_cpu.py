"""Size-constrained k-means, CPU side.

1. The oracle restatement of ``local_utils/sskm_constrained.py`` against ``tests/golden/kmeans_constrained.npz``,
   which ``oracle/gen_golden.py`` produced by running the REAL reference module with only its OR-Tools wrapper
   replaced by the oracle's stand-in solver (the solver itself is parity-unpinned: OR-Tools 9.3.10497 is absent).
2. The product's host solver ``scd_constrained_assign`` (C ABI, host-only - callable without a GPU) against the
   stand-in on the same integer costs: optimal total cost, size bounds, infeasibility, ties, and - at C3 scale - an
   independent optimality certificate (no negative cycle in the residual cluster graph)."""
import os

import numpy as np
import pytest
import torch

from oracle import constrained_oracle as co
from scd_b200 import sskm_constrained as sk

torch.set_num_threads(1)


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, 'kmeans_constrained.npz'))


# ------------------------------------------------------------------------------ oracle vs the reference's outputs
def test_oracle_flow_graph_matches_reference(golden_dir):
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['g_bounds'])
    edges, costs, caps, supplies, n_C, n_X = co.minimum_cost_flow_problem_graph(g['g_X'], g['g_C0'], g['g_D_sqrt'], lo, hi)
    for got, want in ((edges, g['g_edges']), (costs, g['g_costs']), (caps, g['g_caps']), (supplies, g['g_supplies'])):
        assert got.dtype == want.dtype and np.array_equal(got, want)
    assert (n_C, n_X) == (6, 240)
    assert np.array_equal(co.int_costs(g['g_D_sqrt']).reshape(-1), g['g_costs'][:n_X * n_C])


def test_oracle_labels_constrained_matches_reference(golden_dir):
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['g_bounds'])
    labels, inertia = co.labels_constrained(g['g_X'], g['g_C0'], g['g_D_sqrt'], lo, hi)
    assert labels.dtype == np.int32 and np.array_equal(labels, g['g_labels'])
    assert float(inertia) == float(g['g_inertia'])


def test_oracle_test_kmeans_cons_case(golden_dir):
    """The 9 x 2 array of the reference's ``local_utils/test_kmeans_cons.py`` (k=2, sizes 2..5, random_state=0)."""
    g = _g(golden_dir)
    km = co.K_Means(k=2, size_min=2, size_max=5, random_state=0)
    km.fit(torch.from_numpy(g['t9_X']))
    assert np.array_equal(km.labels_.numpy(), g['t9_labels'])
    assert np.array_equal(km.cluster_centers_.numpy(), g['t9_centers'])
    assert float(km.inertia_) == float(g['t9_inertia']) and km.n_iter_ == int(g['t9_n_iter'])


def test_oracle_fit_and_fit_mix_match_reference(golden_dir):
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['g_bounds'])
    km = co.K_Means(k=6, tolerance=1e-4, max_iterations=5, size_min=lo, size_max=hi, init='random', n_init=2, random_state=4,
                    n_jobs=None, pairwise_batch_size=64)
    km.fit(torch.from_numpy(g['g_X']))
    assert km.labels_.dtype == torch.int32 and np.array_equal(km.labels_.numpy(), g['fit_labels'])
    assert np.array_equal(km.cluster_centers_.numpy(), g['fit_centers'])
    assert float(km.inertia_) == float(g['fit_inertia']) and km.n_iter_ == int(g['fit_n_iter'])
    lo, hi = (int(v) for v in g['mix_bounds'])
    km = co.K_Means(k=6, tolerance=1e-4, max_iterations=5, size_min=lo, size_max=hi, init='k-means++', n_init=2, random_state=9,
                    n_jobs=None, pairwise_batch_size=64)
    km.fit_mix(torch.from_numpy(g['mix_u']), torch.from_numpy(g['mix_l']), torch.from_numpy(g['mix_t']))
    assert np.array_equal(km.labels_.numpy(), g['mix_labels'])
    assert np.array_equal(km.cluster_centers_.numpy(), g['mix_centers'])
    assert km.inertia_.item() == float(g['mix_inertia'])
    assert km.n_iter_ == int(g['mix_n_iter']) == len(g['mix_t'])          # the stale-loop-variable quirk (:139)


# ------------------------------------------------------------------------------ the product's host solver
def test_solver_on_the_golden_costs(golden_dir):
    g = _g(golden_dir)
    lo, hi = (int(v) for v in g['g_bounds'])
    cost = g['g_costs'][:240 * 6].reshape(240, 6)
    labels, total, aug = sk.labels_constrained(cost, lo, hi)
    assert total == int(g['g_total_cost']) == int(cost[np.arange(240), labels].sum())
    sizes = np.bincount(labels, minlength=6)
    assert sizes.min() >= lo and sizes.max() <= hi and aug > 0
    assert np.array_equal(labels, g['g_labels'])          # continuous data: the optimum is unique here


def test_solver_matches_standin_on_random_costs():
    rng = np.random.RandomState(0)
    checked = infeasible = 0
    for trial in range(120):
        n, k = int(rng.randint(1, 50)), int(rng.randint(1, 7))
        mode = trial % 3
        if mode == 0:
            cost = rng.randint(0, 2000, (n, k))
        elif mode == 1:
            cost = rng.randint(0, 4, (n, k))                 # massively tied costs
        else:
            x, c = rng.randn(n, 3), rng.randn(k, 3) * 0.3      # crowded centres: the bounds bite
            cost = np.around(np.sqrt(((x[:, None] - c[None]) ** 2).sum(-1)) * 1000).astype(int)
        lo = int(rng.randint(0, n // k + 2))
        hi = int(rng.randint(max(lo, 1), n + 2))
        want = co.optimal_total_cost(cost.astype(np.int32), lo, hi)
        if want is None:
            assert k * lo > n or k * hi < n
            with pytest.raises(Exception, match='min cost flow input'):
                sk.labels_constrained(cost, lo, hi)
            infeasible += 1
            continue
        labels, total, _ = sk.labels_constrained(cost, lo, hi)
        sizes = np.bincount(labels, minlength=k)
        assert sizes.min() >= lo and sizes.max() <= hi
        assert total == want == int(cost[np.arange(n), labels].sum())
        checked += 1
    assert checked > 60 and infeasible > 5


def test_solver_inactive_bounds_is_the_row_argmin():
    rng = np.random.RandomState(1)
    cost = rng.randint(0, 1000, (500, 7)).astype(np.int32)
    labels, total, aug = sk.labels_constrained(cost, 0, 500)
    assert aug == 0 and np.array_equal(labels, cost.argmin(1)) and total == int(cost.min(1).sum())
    labels, _, _ = sk.labels_constrained(np.zeros((0, 3), dtype=np.int32), 0, 5)
    assert labels.shape == (0,)


def _negative_cycle_free(cost, labels, lo, hi):
    """Optimality certificate: the residual graph on the K clusters + T has no negative cycle.  Arc a -> b weighs
    min_{i in a}(cost[i,b] - cost[i,a]) (move the best item of a into b).  A cycle among clusters keeps every size; a
    cycle through T is a path from a donor (size > lo may lose an item) to a receiver (size < hi may gain one).
    Floyd-Warshall over the K clusters covers both."""
    n, k = cost.shape
    sizes = np.bincount(labels, minlength=k)
    big = np.int64(1) << 50
    w = np.full((k, k), big, dtype=np.int64)
    for a in range(k):
        rows = cost[labels == a].astype(np.int64)
        if len(rows):
            w[a] = (rows - rows[:, a:a + 1]).min(axis=0)
        w[a, a] = big
    d = w.copy()
    for m in range(k):
        d = np.minimum(d, d[:, m:m + 1] + d[m:m + 1, :])
    if (np.diag(d) < 0).any():
        return False
    donors, receivers = np.where(sizes > lo)[0], np.where(sizes < hi)[0]
    sub = d[np.ix_(donors, receivers)]
    return not (sub[donors[:, None] != receivers[None, :]] < 0).any()


def test_solver_c3_scale_optimality_certificate():
    """C3 scale (20 000 rows, K = 120) with balanced bounds: thousands of augmentations; optimality is certified
    independently (no improving cycle, no improving donor -> receiver path)."""
    rng = np.random.RandomState(5)
    n, k, d = 20000, 120, 16
    y = (rng.randint(0, k, n) ** 2) % k                       # skewed class sizes
    mu = rng.randn(k, d) * 2
    x = (rng.randn(n, d) + mu[y]).astype(np.float32)
    c = (mu + 0.1 * rng.randn(k, d)).astype(np.float32)
    dist = ((x[:, None, :] - c[None]) ** 2).sum(-1)
    cost = co.int_costs(np.sqrt(dist))
    lo, hi = 150, 180
    labels, total, aug = sk.labels_constrained(cost, lo, hi)
    sizes = np.bincount(labels, minlength=k)
    assert sizes.min() >= lo and sizes.max() <= hi and aug > 1000
    assert total == int(cost[np.arange(n), labels].sum())
    assert _negative_cycle_free(cost, labels, lo, hi)
    # and the certificate does reject a perturbed (feasible but sub-optimal) labelling
    worse = labels.copy()
    a, b = 0, 1
    ia, ib = np.where(labels == a)[0], np.where(labels == b)[0]
    i = ia[np.argmax(cost[ia, b] - cost[ia, a])]
    j = ib[np.argmax(cost[ib, a] - cost[ib, b])]
    worse[i], worse[j] = b, a
    assert not _negative_cycle_free(cost, worse, lo, hi)
