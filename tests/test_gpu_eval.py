"""GPU parity of the evaluation metrics and the device-resident feature split: bit-exact against
the reference-generated fixtures and against the oracle at full size."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle
from scd_b200 import evaluate, features

pytestmark = pytest.mark.gpu


def _cases(g):
    for ci in range(5):
        y, p, mask = g[f'c{ci}_y'], g[f'c{ci}_pred'], g[f'c{ci}_mask']
        names = {c: f'n{100 + c}' for c in range(int(g[f'c{ci}_ncls']))}
        yield ci, y, p, mask, names, [str(x) for x in g[f'c{ci}_cand']]


def test_notebook_known_answer():
    gt = np.array([0] * 5 + [1] * 5 + [2] * 5 + [3] * 5)
    pr = np.array([2] * 4 + [0] * 1 + [1] * 4 + [3] * 1 + [0] * 4 + [3] * 1 + [3] * 5)
    t, o, n, m = evaluate.split_cluster_acc_v2(gt, pr, gt < 2, return_ind_map=True)
    assert (t, o, n, m) == (0.85, 0.8, 0.9, {2: 0, 1: 1, 0: 2, 3: 3})


def test_metrics_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'eval_small.npz'))
    for ci, y, p, mask, names, cand in _cases(g):
        t, o, n, m = evaluate.split_cluster_acc_v2(y, p, mask, return_ind_map=True)
        assert np.array_equal(np.array([t, o, n]), g[f'c{ci}_acc'])                      # float64, bit-exact
        assert np.array_equal(np.array(sorted(m.items())), g[f'c{ci}_map'])
        assert evaluate.split_cluster_acc_v2(torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(mask))[:3] == (t, o, n)
        for sub, sel in (('all', np.ones(len(y), bool)), ('old', mask), ('new', ~mask)):
            assert np.array_equal(np.array(evaluate.evaluate_semantic_acc(y[sel], names, p[sel], cand)), g[f'c{ci}_sem_{sub}'])


@pytest.mark.parametrize('n,n_cls,n_clu', [(127000, 100, 100), (200000, 1000, 1000), (5000, 300, 90), (1, 1, 1)])
def test_contingency_and_metrics_at_scale_against_oracle(n, n_cls, n_clu):
    rng = np.random.RandomState(n % 997)
    y = rng.randint(0, n_cls, size=n)
    pred = np.where(rng.rand(n) < 0.3, rng.randint(0, n_clu, size=n), y % n_clu).astype(np.int64)
    mask = y < max(1, n_cls // 2)
    w, first, colm = evaluate.contingency(pred, y.astype(np.float64), mask=mask)
    assert np.array_equal(w, eval_oracle.contingency(pred, y))
    assert w.sum() == n and np.array_equal(colm, np.bincount(y[mask], minlength=w.shape[0]))   # checksum of checksums
    fo = np.full(w.shape[0], n, dtype=np.int64)
    uy, ui = np.unique(y, return_index=True)
    fo[uy] = ui
    assert np.array_equal(first, fo)
    if n > 1 and n_cls <= 300:                                             # the NumPy Munkres of the oracle is O(D^3)
        assert evaluate.split_cluster_acc_v2(y, pred, mask) == eval_oracle.split_cluster_acc_v2(y, pred, mask)
        names = {c: f'name{c % (n_cls - 3)}' for c in range(n_cls)}        # a few classes share a name
        cand = [f'name{(p * 7) % n_cls}' if p % 5 == 0 else f'name{p}' for p in range(w.shape[0])]
        assert evaluate.evaluate_semantic_acc(y, names, pred, cand) == eval_oracle.evaluate_semantic_acc(y, names, pred, cand)


def test_errors_follow_the_reference():
    with pytest.raises(AssertionError):
        evaluate.split_cluster_acc_v2(np.arange(4), np.arange(3), np.ones(4, bool))
    with pytest.raises(ZeroDivisionError):                                  # no 'new' rows: 0 / 0 on Python ints (:68)
        evaluate.split_cluster_acc_v2(np.arange(4), np.arange(4), np.ones(4, bool))
    with pytest.raises(IndexError):
        evaluate.contingency(np.array([0, 5]), np.array([0, 1]), dim=3)
    with pytest.raises(IndexError):
        evaluate.contingency(np.array([0, -1]), np.array([0, 1]))
    with pytest.raises(ValueError):
        evaluate.evaluate_semantic_acc_ub_lb(np.zeros((2, 8), np.float32), [0, 1], {0: 'a', 1: 'zz'}, ['a', 'b'], torch.zeros(8, 2))


def test_zero_shot_accuracy_matches_the_reference_arithmetic(golden_dir):
    """main_ptsup.py:102-129 on the naming fixture: the reference's accuracy() counts are in the fixture"""
    g = np.load(os.path.join(golden_dir, 'naming_small.npz'))
    n = 700
    feats, W, tgt = g[f'feats_{n}'][:512], torch.from_numpy(g['W']), g[f'acc_tgt_{n}']
    nouns = [f'w{i}' for i in range(W.shape[1])]
    cidx_to_cname = {i: nouns[i] for i in range(W.shape[1])}
    got = evaluate.evaluate_semantic_acc_ub_lb(feats, tgt.astype(np.float64), cidx_to_cname, nouns, W)
    assert got == (float(g[f'acc_{n}'][0]) / 512.0) * 100
    preds = evaluate.get_clip_preds_fast(feats, tgt, cidx_to_cname, nouns, W)
    assert preds.dtype == torch.int64 and preds.is_cuda
    val = g[f'ptsup_val_{n}'][:512]
    clear = (val[:, 0] - val[:, 1]) > 1e-3                                  # top-1/top-2 margin (x100 logits), as in test_gpu_naming
    assert clear.mean() > 0.97 and np.array_equal(preds.cpu().numpy()[clear], g[f'ptsup_idx_{n}'][:512, 0][clear])


def test_feature_set_splits_like_the_driver(golden_dir):
    d = features.load_features(os.path.join(golden_dir, 'features_all.pt'))
    fs = features.FeatureSet(d)
    lab = d['mask_lab']
    assert torch.equal(fs.l_feats.cpu(), torch.from_numpy(d['all_feats'][lab]))          # main_unsup.py:323-326
    assert torch.equal(fs.u_feats.cpu(), torch.from_numpy(d['all_feats'][~lab]))
    assert fs.l_targets.dtype == torch.float64 and np.array_equal(fs.l_targets.cpu().numpy(), d['targets'][lab])
    assert np.array_equal(fs.u_targets.cpu().numpy(), d['targets'][~lab])
    assert np.array_equal(fs.mask, d['mask_cls'][~lab].astype(bool))                    # :329-331
    assert fs.bf16().dtype == torch.bfloat16 and torch.equal(fs.bf16().float().cpu(), fs.u_feats.cpu().bfloat16().float())
