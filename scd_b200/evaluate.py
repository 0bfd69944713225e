"""Evaluation metrics either side of the naming round, same signatures as the reference.

Reference functions (paths relative to the reference checkout):
  * ``split_cluster_acc_v2(y_true, y_pred, mask, return_ind_map=False)``
        ``gcd/project_utils/cluster_and_log_utils.py:28-76`` (called ``main_unsup.py:376,564,617``, ``main_ptsup.py:395,580,621,680``)
  * ``evaluate_semantic_acc(u_targets, cidx_to_cname, u_preds, cand_names)``
        ``main_unsup.py:149-167`` (copy ``main_ptsup.py:168-186``; called once per voting round, ``:620-624``)
  * ``get_clip_preds_fast`` / ``evaluate_semantic_acc_ub_lb``  ``main_ptsup.py:78-99`` / ``:102-129``

The reference fills the contingency matrix with a Python loop over all rows (``w[y_pred[i], y_true[i]] += 1``,
O(N) interpreter steps per call, three calls per voting round) and scans Python lists of strings per row.  Here one
kernel (``scd_contingency``) builds ``w`` and the first-occurrence order of the classes from the device-resident label
vectors; everything after it works on the ``D x D`` matrix (D = number of clusters / classes): the Hungarian matching
(``scd_linear_assignment``, the reference's tie-breaking) and exact integer sums.  Results are the reference's values
bit for bit (same integer counts, same float64 divisions, same summation order).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, naming


def _label_tensor(y):
    """Device int64 or float64 vector (float labels are truncated toward zero in the kernel = ``.astype(int)``)."""
    if not torch.is_tensor(y):
        y = torch.from_numpy(np.ascontiguousarray(np.asarray(y)))
    if y.dtype == torch.bool:
        y = y.to(torch.int64)
    if y.dtype not in (torch.int64, torch.float64):
        y = y.to(torch.float64 if y.dtype.is_floating_point else torch.int64)
    return y.to('cuda').contiguous().view(-1)


def contingency(y_pred, y_true, dim: int | None = None, mask=None):
    """``w[y_pred[i], y_true[i]] += 1`` (``cluster_and_log_utils.py:45-49``) on the device.

    Returns ``(w [D, D] int64, first_row [D] int64, col_masked [D] int64)`` as ndarrays; ``first_row[t]`` is the first
    row whose true class is ``t`` (``N`` if absent), ``col_masked[t]`` the number of rows of class ``t`` with ``mask``
    set (zeros without a mask).  ``dim`` defaults to ``max(y_pred.max(), y_true.max()) + 1`` like the reference."""
    naming._require_cuda()
    p, t = _label_tensor(y_pred), _label_tensor(y_true)
    if p.numel() != t.numel():
        raise AssertionError('y_pred.size != y_true.size')               # the reference's assert, :46
    n = int(p.numel())
    m = None
    if mask is not None:
        m = mask if torch.is_tensor(mask) else torch.from_numpy(np.ascontiguousarray(np.asarray(mask)))
        m = m.to('cuda').ne(0).to(torch.uint8).contiguous().view(-1)
        if m.numel() != n:
            raise IndexError(f'boolean index did not match indexed array: mask has {m.numel()} entries, labels {n}')
    if dim is None:
        if n == 0:
            raise ValueError('zero-size array to reduction operation maximum which has no identity')    # numpy's error
        dim = int(max(int(p.max().item()), int(t.max().item()))) + 1
    w = torch.empty(dim, dim, dtype=torch.int64, device='cuda')
    first = torch.empty(dim, dtype=torch.int64, device='cuda')
    colm = torch.zeros(dim, dtype=torch.int64, device='cuda')
    bad = torch.zeros(1, dtype=torch.int32, device='cuda')
    lib = _lib.load()
    _lib.check(lib.scd_contingency(p.data_ptr(), int(p.dtype == torch.float64), t.data_ptr(), int(t.dtype == torch.float64),
                                   n, dim, _lib.ptr(m), w.data_ptr(), first.data_ptr(), colm.data_ptr(), bad.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), 'scd_contingency')
    if int(bad.item()):
        raise IndexError(f'label outside [0, {dim}) (the reference would index out of bounds or wrap)')
    return w.cpu().numpy(), first.cpu().numpy(), colm.cpu().numpy()


def _cluster_acc_from_contingency(w, col_old, return_ind_map=False):
    """Host half of ``split_cluster_acc_v2`` (``cluster_and_log_utils.py:52-76``) on the ``D x D`` contingency matrix and
    the per-class counts of rows under the old/new mask: Hungarian matching with the reference's tie-breaking, then
    the reference's integer sums and float64 divisions."""
    n = int(w.sum())
    col_all = w.sum(axis=0)
    old_classes = np.nonzero(col_old)[0]                                   # set(y_true[mask]), :43
    new_classes = np.nonzero(col_all - col_old)[0]                         # set(y_true[~mask]), :44
    ind = naming.linear_assignment(w.max() - w)                            # :52
    ind_map = {int(j): int(i) for i, j in ind}                             # :53  true class -> cluster
    total_acc = sum(w[i, j] for i, j in ind) * 1.0 / n                     # :54
    accs = []
    for classes in (old_classes, new_classes):                             # :56-68
        hit = 0
        inst = 0
        for c in classes:
            hit += w[ind_map[int(c)], c]
            inst += col_all[c]
        accs.append(hit / inst)            # an empty side is 0 / 0 -> ZeroDivisionError, like the reference's int / int
    if return_ind_map:
        return total_acc, accs[0], accs[1], ind_map
    return total_acc, accs[0], accs[1]


def split_cluster_acc_v2(y_true, y_pred, mask, return_ind_map=False):
    """Clustering accuracy after one Hungarian matching on all rows; 'old' = classes met under ``mask``, 'new' =
    classes met under ``~mask`` (``cluster_and_log_utils.py:28-76``).  Returns ``(total_acc, old_acc, new_acc[, ind_map])``."""
    w, _, col_old = contingency(y_pred, y_true, mask=mask)
    return _cluster_acc_from_contingency(w, col_old, return_ind_map)


def _semantic_acc_from_contingency(w, first, cidx_to_cname, cand_names):
    """Host half of ``evaluate_semantic_acc`` (``main_unsup.py:149-167``): names are compared once per (cluster, class)
    cell; the per-name accuracies are summed in the order the class names are first met in the rows (``first``)."""
    n = int(w.sum())
    col_all = w.sum(axis=0)
    classes = [int(c) for c in np.argsort(first, kind='stable') if col_all[c] > 0]      # order of first appearance (:152)
    per_name = {}
    matched_all = 0
    for c in classes:
        name = cidx_to_cname[c]
        hit = 0
        for p in np.nonzero(w[:, c])[0]:
            if cand_names[int(p)] == name:
                hit += int(w[p, c])
        rec = per_name.setdefault(name, [0, 0])
        rec[0] += hit
        rec[1] += int(col_all[c])
        matched_all += hit
    acc = {name: hit / float(cnt) for name, (hit, cnt) in per_name.items()}             # :160-163
    return float(sum(acc.values())) / len(acc.values()), matched_all / float(n)         # :165-167


def evaluate_semantic_acc(u_targets, cidx_to_cname, u_preds, cand_names):
    """``main_unsup.py:149-167``: ``(semantic_acc_avg, semantic_acc_all)`` - a row matches when the name of its class
    equals the name voted for its cluster.  The per-row string compares of the reference become one contingency
    matrix (``scd_contingency``)."""
    w, first, _ = contingency(u_preds, u_targets)
    return _semantic_acc_from_contingency(w, first, cidx_to_cname, cand_names)


def _target_name_idx(targets, cidx_to_cname, nouns):
    """``[nouns.index(cidx_to_cname[t]) for t in targets]`` (``main_ptsup.py:89,113``) with one dict instead of an
    O(V) list scan per row (first occurrence wins, like ``list.index``)."""
    pos = {}
    for i, name in enumerate(nouns):
        pos.setdefault(name, i)
    t = np.asarray(targets.cpu() if torch.is_tensor(targets) else targets)
    try:
        return np.array([pos[cidx_to_cname[x]] for x in t.tolist()], dtype=np.int64)
    except KeyError as e:
        raise ValueError(f'{e.args[0]!r} is not in list') from None        # list.index's error


def get_clip_preds_fast(clip_feats, targets, cidx_to_cname, nouns, zeroshot_weights):
    """``main_ptsup.py:78-99``: ``argmax(100 * feats @ zeroshot_weights)`` over the whole vocabulary, ``LongTensor[n]``
    on the device.  ``targets`` are only validated (the reference maps them to noun indices and drops the result)."""
    _target_name_idx(targets, cidx_to_cname, nouns)
    return naming.clip_preds(clip_feats, zeroshot_weights)


def evaluate_semantic_acc_ub_lb(clip_feats, targets, cidx_to_cname, nouns, zeroshot_weights):
    """``main_ptsup.py:102-129``: top-1 zero-shot accuracy in percent (top-5 is computed and dropped there too)."""
    tgt = _target_name_idx(targets, cidx_to_cname, nouns)
    n = int(clip_feats.shape[0])
    top1, _top5 = naming.accuracy(clip_feats, zeroshot_weights, tgt, topk=(1, 5))
    return (top1 / float(n)) * 100
