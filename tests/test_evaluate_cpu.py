"""CPU only: the host halves of the evaluation metrics (everything after the contingency kernel: Hungarian matching
through the C ABI, integer sums, float64 divisions, first-appearance summation order) against the fixtures produced
by the reference's own functions.  The contingency matrix itself comes from NumPy here; the kernel that builds it on
the device is checked in tests/test_gpu_eval.py."""
import os

import numpy as np
import pytest

from scd_b200 import evaluate


def _contingency(pred, y, mask=None):
    pred, y = np.asarray(pred).astype(int), np.asarray(y).astype(int)
    d = int(max(pred.max(), y.max())) + 1
    w = np.zeros((d, d), dtype=np.int64)
    np.add.at(w, (pred, y), 1)
    first = np.full(d, len(y), dtype=np.int64)
    uy, ui = np.unique(y, return_index=True)
    first[uy] = ui
    colm = np.bincount(y[np.asarray(mask, bool)], minlength=d).astype(np.int64) if mask is not None else np.zeros(d, np.int64)
    return w, first, colm


def test_notebook_known_answer_through_the_host_half():
    gt = np.array([0] * 5 + [1] * 5 + [2] * 5 + [3] * 5)
    pr = np.array([2] * 4 + [0] * 1 + [1] * 4 + [3] * 1 + [0] * 4 + [3] * 1 + [3] * 5)
    w, _, colm = _contingency(pr, gt, gt < 2)
    assert evaluate._cluster_acc_from_contingency(w, colm, True) == (0.85, 0.8, 0.9, {2: 0, 1: 1, 0: 2, 3: 3})


def test_host_halves_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'eval_small.npz'))
    for ci in range(5):
        y, p, mask = g[f'c{ci}_y'], g[f'c{ci}_pred'], g[f'c{ci}_mask']
        w, first, colm = _contingency(p, y, mask)
        t, o, n, m = evaluate._cluster_acc_from_contingency(w, colm, True)
        assert np.array_equal(np.array([t, o, n]), g[f'c{ci}_acc'])                      # float64, bit-exact
        assert np.array_equal(np.array(sorted(m.items())), g[f'c{ci}_map'])
        names = {c: f'n{100 + c}' for c in range(int(g[f'c{ci}_ncls']))}
        cand = [str(x) for x in g[f'c{ci}_cand']]
        for sub, sel in (('all', np.ones(len(y), bool)), ('old', mask), ('new', ~mask)):
            ws, fs, _ = _contingency(p[sel], y[sel])
            got = evaluate._semantic_acc_from_contingency(ws, fs, names, cand)
            assert np.array_equal(np.array(got), g[f'c{ci}_sem_{sub}'])


def test_empty_side_raises_like_the_reference():
    w, _, colm = _contingency(np.arange(4), np.arange(4), np.ones(4, bool))
    with pytest.raises(ZeroDivisionError):                                   # no 'new' rows: 0 / 0 on Python ints (:68)
        evaluate._cluster_acc_from_contingency(w, colm)


def test_target_name_lookup_follows_list_index():
    nouns = ['a', 'b', 'a', 'c']
    assert evaluate._target_name_idx(np.array([0., 1., 2.]), {0: 'a', 1: 'c', 2: 'b'}, nouns).tolist() == [0, 3, 1]
    with pytest.raises(ValueError):
        evaluate._target_name_idx([0], {0: 'zz'}, nouns)
