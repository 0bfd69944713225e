"""CPU restatement of the evaluation metrics either side of the naming round.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned by ``tests/golden/eval_small.npz``, which
``oracle/gen_golden.py`` writes by running the real reference functions:
  * ``split_cluster_acc_v2``  ``gcd/project_utils/cluster_and_log_utils.py:28-76``
  * ``evaluate_semantic_acc`` ``main_unsup.py:149-167`` (identical copy ``main_ptsup.py:168-186``)
and by the reference's one known-answer test for this code, ``gcd/notebooks/demo_acc_v2.ipynb``
(``0.85 0.8 0.9 {2: 0, 1: 1, 0: 2, 3: 3}``).
"""
from __future__ import annotations

import numpy as np

from . import hungarian_oracle


def contingency(y_pred, y_true, dim=None):
    """``cluster_and_log_utils.py:45-49``: ``w[y_pred[i], y_true[i]] += 1`` over all rows, ``D = max(max) + 1``."""
    y_pred = np.asarray(y_pred).astype(int)
    y_true = np.asarray(y_true).astype(int)
    if dim is None:
        dim = int(max(y_pred.max(), y_true.max())) + 1
    w = np.zeros((dim, dim), dtype=int)
    np.add.at(w, (y_pred, y_true), 1)
    return w


def split_cluster_acc_v2(y_true, y_pred, mask, return_ind_map=False):
    """``cluster_and_log_utils.py:28-76``: one Hungarian matching on all rows, then accuracy on the classes met
    under ``mask`` ('old') and under ``~mask`` ('new')."""
    y_true = np.asarray(y_true).astype(int)                                # :41
    y_pred = np.asarray(y_pred)
    mask = np.asarray(mask)
    old_gt, new_gt = set(y_true[mask]), set(y_true[~mask])                 # :43-44
    assert y_pred.size == y_true.size                                      # :46
    w = contingency(y_pred, y_true)                                        # :47-50
    ind = hungarian_oracle.linear_assignment(w.max() - w)                  # :52
    ind_map = {j: i for i, j in ind}                                       # :53  true class -> cluster
    total_acc = sum(w[i, j] for i, j in ind) * 1.0 / y_pred.size           # :54
    out = [total_acc]
    for classes in (old_gt, new_gt):                                       # :56-68
        hit = 0
        inst = 0
        for c in classes:
            hit += w[ind_map[c], c]
            inst += sum(w[:, c])
        out.append(hit / inst)
    if return_ind_map:
        return out[0], out[1], out[2], ind_map
    return out[0], out[1], out[2]


def evaluate_semantic_acc(u_targets, cidx_to_cname, u_preds, cand_names):
    """``main_unsup.py:149-167``: a row matches when its class's name equals the name voted for its cluster.
    Returns ``(mean over class names of the per-name accuracy, accuracy over all rows)``; the per-name
    accuracies are summed in the order the names are first met (dict insertion order)."""
    per_name = {}
    matched = 0
    n = 0
    for t, p in zip(u_targets, u_preds):                                   # :152-158
        name = cidx_to_cname[t]
        ok = 1 if name == cand_names[p] else 0
        rec = per_name.setdefault(name, [0, 0])
        rec[0] += ok
        rec[1] += 1
        matched += ok
        n += 1
    acc = {name: hit / float(cnt) for name, (hit, cnt) in per_name.items()}     # :160-163
    return float(sum(acc.values())) / len(acc), matched / float(n)              # :165-167
