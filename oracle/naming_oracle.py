"""Oracle: image x vocabulary scoring, per-image top-k, per-cluster voting, name assignment and
re-assignment of SCD, restated on torch-CPU / NumPy / ``collections.Counter``.

TEST INFRASTRUCTURE - not shipped, never imported by ``scd_b200``.

The scoring block and the voting loop are *inline script code* in the reference (not importable
functions), so these are line-for-line restatements of:
  * ``main_unsup.py:504-531``  scoring + softmax + top-k (unsupervised driver)
  * ``main_ptsup.py:526-545``  scoring + top-k, no softmax (partially supervised driver)
  * ``main_ptsup.py:78-99``    ``get_clip_preds_fast`` (argmax over the whole vocabulary)
  * ``main_unsup.py:568-614``  vote -> assign_name -> narrow vocabulary -> re-assign loop
  * ``main_ptsup.py:629-676``  same with labelled-class filtering and the index-space quirk
  * ``local_utils/clip_lang_util.py:151-154`` ``accuracy``; ``:156-180`` ``assign_name``
``assign_name`` and ``accuracy`` are importable from the reference; ``oracle/gen_golden.py`` calls
the real ones to generate ``tests/golden/naming_*.npz`` which pin this file.
"""
from __future__ import annotations

import copy
from collections import Counter

import numpy as np
import torch
import torch.nn.functional as F

from .hungarian_oracle import linear_assignment

BATCH_ROWS = 1024        # ``batch_feat_size`` main_unsup.py:507 / main_ptsup.py:529


def _row_batches(n: int):
    """The reference walks ``int(n / 1024) + 1`` batches (main_unsup.py:508-515): the last one is
    ragged, and it is *empty* when 1024 divides n (it still runs and concatenates a [0,k] block)."""
    for b in range(int(n / BATCH_ROWS) + 1):
        lo = b * BATCH_ROWS
        hi = n if (b + 1) * BATCH_ROWS > n else (b + 1) * BATCH_ROWS
        yield lo, hi


def score_topk(feats, zeroshot_weights: torch.Tensor, k: int, variant: str):
    """``variant='unsup'``: main_unsup.py:504-531 - ``100. * (feat @ W)`` then ``F.softmax``;
    ``variant='ptsup'``: main_ptsup.py:526-545 - ``100. * feat @ W`` (= ``(100*feat) @ W``: the scale is
    applied to the features *before* the contraction there, which rounds differently), no softmax.

    feats [N,D] tensor or ndarray, zeroshot_weights [D,V].  Returns (indices [N,k] int64,
    values [N,k]) - ``topk(k, 1, True, True)`` i.e. largest, sorted.
    """
    assert variant in ('unsup', 'ptsup')
    idx_parts, val_parts = [], []
    for lo, hi in _row_batches(feats.shape[0]):
        batch = feats[lo:hi]
        if not torch.is_tensor(batch):
            batch = torch.from_numpy(batch)                     # ref :522 (.cuda() there)
        if variant == 'unsup':
            logits = 100. * (batch @ zeroshot_weights)          # main_unsup.py:519/:524
            logits = F.softmax(logits, dim=1)                   # main_unsup.py:527 (implicit dim -> 1 for 2-D)
        else:
            logits = 100. * batch @ zeroshot_weights            # main_ptsup.py:538/:540
        top = logits.topk(k, 1, True, True)                     # ref :528-529 (called twice there)
        idx_parts.append(top[1])
        val_parts.append(top[0])
    return torch.cat(idx_parts, dim=0), torch.cat(val_parts, dim=0)


def clip_preds(feats, zeroshot_weights: torch.Tensor) -> torch.Tensor:
    """main_ptsup.py:78-99 ``get_clip_preds_fast`` without the target bookkeeping: batched
    ``argmax(100 * feats @ W)`` over the whole vocabulary.  Note the association there is
    ``(100. * feats) @ W`` (:92) - kept."""
    out = []
    for lo, hi in _row_batches(feats.shape[0]):
        batch = feats[lo:hi]
        if not torch.is_tensor(batch):
            batch = torch.from_numpy(batch)
        logits = 100. * batch @ zeroshot_weights
        out.append(logits.argmax(dim=-1).view(-1))
    return torch.cat(out, dim=0)


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """local_utils/clip_lang_util.py:151-154 - number (not fraction) of rows whose target is in the top-k."""
    pred = output.topk(max(topk), 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [float(correct[:kk].reshape(-1).float().sum(0, keepdim=True).cpu().numpy()[0]) for kk in topk]


def reassign(feats, w_selected: torch.Tensor) -> np.ndarray:
    """main_unsup.py:604-614 / main_ptsup.py:671-676: ``argmax(100 * feats @ W_sel, -1)`` -> NumPy int64."""
    if not torch.is_tensor(feats):
        feats = torch.from_numpy(feats)
    logits = 100. * (feats @ w_selected)
    return logits.argmax(dim=-1).view(-1).cpu().numpy()


def vote(name_idx_topk: torch.Tensor, u_preds: np.ndarray, cluster_ids, top_k: int,
         known_name_idx=None):
    """main_unsup.py:575-577 / main_ptsup.py:636-638: one ``Counter`` per voting cluster over the
    row-major flattened ``name_idx_topk[u_preds == i, :top_k]``; ptsup drops ``known_name_idx``."""
    cluster_to_counter = {}
    for i in cluster_ids:
        flat = name_idx_topk[u_preds == i, :top_k].reshape(-1).cpu().numpy()
        if known_name_idx is None:
            cluster_to_counter[i] = Counter(x for x in flat)
        else:
            cluster_to_counter[i] = Counter(x for x in flat if x not in known_name_idx)
    return cluster_to_counter


def voted_candidates(cluster_to_counter, cluster_ids, num_common_vote: int):
    """main_unsup.py:579-586 / main_ptsup.py:640-648: union of every cluster's
    ``most_common(num_common_vote)`` names, then ``list(set(...))`` (CPython set order)."""
    names = []
    for i in cluster_ids:
        for name, _count in cluster_to_counter[i].most_common(num_common_vote):
            names += [name]
    return list(set(names))


def assign_name(unique_name_idx, cluster_to_counter, num_common=4):
    """local_utils/clip_lang_util.py:156-180: square int matrix ``w`` (clusters x candidate names)
    filled from ``most_common(num_common)``, then Hungarian on ``w.max() - w``."""
    col_of = {name: j for j, name in enumerate(unique_name_idx)}
    clusters = list(cluster_to_counter.keys())
    dim = max(len(unique_name_idx), len(clusters))
    w = np.zeros((dim, dim), dtype=int)
    for row, cid in enumerate(clusters):
        for name, cnt in cluster_to_counter[cid].most_common(num_common):
            w[row, col_of[name]] += cnt
    ind = linear_assignment(w.max() - w)
    return ind, w


def naming_loop_unsup(name_idx_topk: torch.Tensor, u_preds: np.ndarray, clip_u_feats,
                      zeroshot_weights: torch.Tensor, n_cluster: int, top_k=5,
                      num_common_vote=20, num_common_linear=4, max_rounds=50):
    """main_unsup.py:568-614 with names represented by their vocabulary index (``nouns[i]`` <-> ``i``,
    so ``nouns.index(n)`` :601 is the identity).  Returns the per-round trace."""
    cur, prev = [-1], [-2]                 # ref :560-561 ([0] / [1] there: any two different sets)
    trace = []
    while set(cur) != set(prev) and len(trace) < max_rounds:
        cluster_ids = list(set(u_preds))                                           # ref :573
        c2c = vote(name_idx_topk, u_preds, cluster_ids, top_k)                     # ref :575-577
        uniq = voted_candidates(c2c, cluster_ids, num_common_vote)                 # ref :579-586
        ind, w = assign_name(uniq, c2c, num_common=num_common_linear)              # ref :588
        prev = copy.deepcopy(cur)
        cur = [int(uniq[x[1]]) for x in ind[:n_cluster]]                           # ref :594
        w_sel = torch.stack([zeroshot_weights[:, n] for n in cur], dim=1)          # ref :601-602
        u_preds = reassign(clip_u_feats, w_sel)                                    # ref :604-614
        trace.append(dict(voted=list(cur), u_preds=u_preds.copy(), n_unique=len(uniq)))
    return trace


def naming_loop_ptsup(name_idx_topk: torch.Tensor, all_preds: np.ndarray, mask_lab: np.ndarray,
                      clip_u_feats, zeroshot_weights: torch.Tensor, lab_name_idx, n_cluster: int,
                      top_k=5, num_common_vote=20, num_common_linear=4, max_rounds=50, nouns=None):
    """main_ptsup.py:588-676.  ``nouns=None``: names are vocabulary indices (``sorted(cand_names)`` :659 sorts
    integers - the committed fixtures use index-named vocabularies on both sides).  ``nouns`` = the vocabulary's
    name strings: the loop works on strings exactly as the reference does - ``sorted`` is lexicographic,
    ``nouns.index`` resolves duplicates to the first occurrence (:601, :669), the convergence test compares name sets,
    and ``list(set(cand_names) - set(lab_names))`` (:664) takes CPython's string-set order of the running process.

    Reproduces the reference's index-space quirk: from round 2 on ``unlab_cluster_idx`` and
    ``known_name_idx`` are positions in ``cand_names`` (:662-666) while ``name_idx_topk`` still holds
    vocabulary indices."""
    u_preds = all_preds[~mask_lab]                                                 # ref :592
    l_preds = all_preds[mask_lab]                                                  # ref :593
    name_of = (lambda i: i) if nouns is None else (lambda i: nouns[i])
    first = None if nouns is None else {}
    if nouns is not None:
        for i, s in enumerate(nouns):
            first.setdefault(s, i)
    index_of = (lambda s: s) if nouns is None else (lambda s: first[s])            # nouns.index(s)
    lab_names = [name_of(int(i)) for i in lab_name_idx]                            # ref :598
    num_unlab = n_cluster - len(lab_names)                                         # ref :602
    known = [index_of(s) for s in lab_names]                                       # ref :603 (index in nouns)
    unlab_clusters = list(set(set(all_preds)) - set(l_preds))                      # ref :625
    cur, prev = [-1], [-2]
    trace = []
    while set(cur) != set(prev) and len(trace) < max_rounds:
        c2c = vote(name_idx_topk, u_preds, unlab_clusters, top_k, known_name_idx=known)   # ref :636-638
        uniq = voted_candidates(c2c, unlab_clusters, num_common_vote)                     # ref :640-648
        ind, w = assign_name(uniq, c2c, num_common=num_common_linear)                     # ref :649
        prev = copy.deepcopy(cur)
        cur = [name_of(int(uniq[x[1]])) for x in ind[:num_unlab]]                         # ref :655
        cand = sorted(set(cur + lab_names))                                               # ref :657-659
        lab_class_index = [cand.index(n) for n in lab_names]                              # ref :662
        unlab_clusters = [cand.index(n) for n in list(set(cand) - set(lab_names))]        # ref :664
        known = copy.deepcopy(lab_class_index)                                            # ref :666
        cand_idx = [index_of(s) for s in cand]
        w_sel = torch.stack([zeroshot_weights[:, n] for n in cand_idx], dim=1)            # ref :668-669
        u_preds = reassign(clip_u_feats, w_sel)                                           # ref :671-676
        trace.append(dict(voted=list(cur), cand=list(cand), u_preds=u_preds.copy(), n_unique=len(uniq)))
    return trace
