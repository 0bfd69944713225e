"""Oracle: (semi-supervised) k-means of SCD, restated on torch-CPU.

TEST INFRASTRUCTURE - not shipped, never imported by ``scd_b200``.

Follows (paths relative to the reference checkout):
  * ``local_utils/faster_mix_k_means_pytorch.py``            (``K_Means`` :8-175, ``pairwise_distance`` :177-212)
  * ``gcd/methods/clustering/faster_mix_k_means_pytorch.py`` (same arithmetic; ``mode=`` kwarg :49-59,
    guarded ``kpp`` :82-110) - the copy ``main_unsup.py:25`` / ``main_ptsup.py:25`` import.

Pinned by ``tests/golden/kmeans_*.npz`` (generated from the real reference by
``oracle/gen_golden.py``); on CPU the restatement issues the same torch ops in the same
order, so it is expected to be bit-identical to the reference there.
"""
from __future__ import annotations

import numpy as np
import torch
from sklearn.utils import check_random_state


# --------------------------------------------------------------------------------------
# a1  pairwise_distance   (local_utils/faster_mix_k_means_pytorch.py:177-212)
# --------------------------------------------------------------------------------------
def pairwise_distance(data1: torch.Tensor, data2: torch.Tensor, batch_size=None) -> torch.Tensor:
    """Squared Euclidean distance, direct form ``sum_d (x_d - c_d)^2``.

    ref :185-188 broadcast views, :190-194 un-batched branch, :195-210 batched branch whose
    result buffer is ``torch.zeros(N, K)`` - i.e. default dtype, **CPU**, whatever the inputs.
    """
    lhs = data1.unsqueeze(1)            # [N,1,D]   ref :186
    rhs = data2.unsqueeze(0)            # [1,K,D]   ref :189
    if batch_size is None:              # ref :191-194
        return ((lhs - rhs) ** 2).sum(dim=-1)
    n = data1.shape[0]
    out = torch.zeros(n, data2.shape[0])            # ref :197 (CPU, fp32)
    for lo in range(0, n, batch_size):              # ref :198-210 (the while/if/elif walks the same slices)
        hi = min(lo + batch_size, n)
        out[lo:hi] = ((lhs[lo:hi] - rhs) ** 2).sum(dim=-1)
    return out


# --------------------------------------------------------------------------------------
# a2 / a3 / a4  one Lloyd iteration pieces
# --------------------------------------------------------------------------------------
def estep(X: torch.Tensor, centers: torch.Tensor, batch_size=None):
    """ref :58-60 - distances, row min/argmin, inertia = sum of the minima."""
    dist = pairwise_distance(X, centers, batch_size)
    mindist, labels = torch.min(dist, dim=1)
    return labels, mindist, mindist.sum()


def mstep(X: torch.Tensor, labels: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """ref :61-64 - ``centers[j] = mean(X[labels == j])`` in place; an empty cluster yields a NaN row
    (mean over zero rows), there is no relocation."""
    for j in range(centers.shape[0]):
        members = torch.nonzero(labels == j).squeeze()
        centers[j] = torch.index_select(X, 0, members).mean(dim=0)
    return centers


def center_shift(centers: torch.Tensor, centers_old: torch.Tensor) -> torch.Tensor:
    """ref :71 - sum over clusters of the L2 norm of the move (the caller squares it, :72)."""
    return torch.sum(torch.sqrt(torch.sum((centers - centers_old) ** 2, dim=1)))


class K_Means:
    """Restatement of the reference class; constructor per ref :9-17 (+ ``mode`` of the gcd copy :49-59)."""

    def __init__(self, k=3, tolerance=1e-4, max_iterations=100, init='k-means++', n_init=10,
                 random_state=None, n_jobs=None, pairwise_batch_size=None, mode=None,
                 guarded_kpp=False):
        self.k = k
        self.tolerance = tolerance
        self.max_iterations = max_iterations
        self.init = init
        self.n_init = n_init
        self.random_state = random_state
        self.n_jobs = n_jobs
        self.pairwise_batch_size = pairwise_batch_size
        self.mode = mode
        # gcd copy :104-107 keeps the previous index when no cum_prob >= r; local copy :34 raises IndexError
        self.guarded_kpp = guarded_kpp

    # a5  k-means++ seeding, ref :20-36
    def kpp(self, X, pre_centers=None, k=10, random_state=None):
        rs = check_random_state(random_state)
        if pre_centers is not None:
            C = pre_centers
        else:
            C = X[rs.randint(0, len(X))]                       # ref :25
        C = C.view(-1, X.shape[1])
        ind = None
        while C.shape[0] < k:
            dist = pairwise_distance(X, C, self.pairwise_batch_size).view(-1, C.shape[0])  # ref :28-29
            d2, _ = torch.min(dist, dim=1)                     # ref :30
            prob = d2 / d2.sum()                               # ref :31
            cum_prob = torch.cumsum(prob, dim=0)               # ref :32
            r = rs.rand()                                      # ref :33  one host draw per added centre
            hits = (cum_prob >= r).nonzero()
            if self.guarded_kpp and len(hits) == 0:
                pass                                           # gcd :104-105 - reuse previous ``ind``
            else:
                ind = hits[0][0]                               # ref :34 (IndexError when empty)
            C = torch.cat((C, X[ind].view(1, -1)), dim=0)      # ref :35
        return C

    # a6  unsupervised restart, ref :39-75
    def fit_once(self, X, random_state):
        centers = torch.zeros(self.k, X.shape[1]).type_as(X)
        if self.init == 'k-means++':
            centers = self.kpp(X, k=self.k, random_state=random_state)
        elif self.init == 'random':
            rs = check_random_state(self.random_state)         # ref :46 - note: self.random_state, not the arg
            idx = rs.choice(len(X), self.k, replace=False)
            for i in range(self.k):
                centers[i] = X[idx[i]]
        else:
            for i in range(self.k):
                centers[i] = X[i]
        best_labels = best_inertia = best_centers = None
        n_done = 0
        for it in range(self.max_iterations):
            n_done = it + 1
            centers_old = centers.clone()
            labels, _, inertia = estep(X, centers, self.pairwise_batch_size)
            mstep(X, labels, centers)
            if best_inertia is None or inertia < best_inertia:  # ref :66-69
                best_labels, best_centers, best_inertia = labels.clone(), centers.clone(), inertia
            if center_shift(centers, centers_old) ** 2 < self.tolerance:   # ref :71-74
                break
        return best_labels, best_inertia, best_centers, n_done   # ref :75 (i + 1)

    # a6  semi-supervised restart, ref :77-127
    def fit_mix_once(self, u_feats, l_feats, l_targets, random_state):
        l_classes = torch.unique(l_targets)                                      # ref :80 sorted unique
        l_centers = torch.stack([l_feats[l_targets.eq(c).nonzero().squeeze(1)].mean(0)
                                 for c in l_classes])                            # ref :78-82
        cat_feats = torch.cat((l_feats, u_feats))                                # ref :83
        labels = -torch.ones(len(cat_feats)).type_as(cat_feats).long()           # ref :88
        classes_np = l_classes.cpu().long().numpy()
        targets_np = l_targets.cpu().long().numpy()
        l_num = len(targets_np)
        remap = {cid: new for new, cid in enumerate(classes_np)}                 # ref :93
        i = None
        for i in range(l_num):                                                   # ref :94-95
            labels[i] = remap[targets_np[i]]
        centers = self.kpp(u_feats, l_centers, k=self.k, random_state=random_state)  # ref :98
        best_labels = best_inertia = best_centers = None
        for _it in range(self.max_iterations):
            centers_old = centers.clone()
            u_labels, _, u_inertia = estep(u_feats, centers, self.pairwise_batch_size)   # ref :105-107
            l_mindist = torch.sum((l_feats - centers[labels[:l_num]]) ** 2, dim=1)       # ref :108
            inertia = u_inertia + l_mindist.sum()                                        # ref :109-110
            labels[l_num:] = u_labels                                                    # ref :111
            mstep(cat_feats, labels, centers)                                            # ref :113-116
            if best_inertia is None or inertia < best_inertia:                           # ref :118-121
                best_labels, best_centers, best_inertia = labels.clone(), centers.clone(), inertia
            if center_shift(centers, centers_old) ** 2 < self.tolerance:                 # ref :123-126
                break
        # ref :127 returns ``i + 1`` where ``i`` is the stale loop variable of :94 => n_iter == l_num
        return best_labels, best_inertia, best_centers, i + 1

    def fit(self, X):                                                    # ref :129-140 (n_jobs == 1 branch)
        rs = check_random_state(self.random_state)
        best = None
        for _ in range(self.n_init):
            labels, inertia, centers, n_iters = self.fit_once(X, rs)
            if best is None or inertia < best:
                self.labels_, self.cluster_centers_ = labels.clone(), centers.clone()
                best = inertia
                self.inertia_, self.n_iter_ = inertia, n_iters

    def fit_mix(self, u_feats, l_feats, l_targets):                       # ref :153-164
        rs = check_random_state(self.random_state)
        best = None
        for _ in range(self.n_init):
            labels, inertia, centers, n_iters = self.fit_mix_once(u_feats, l_feats, l_targets, rs)
            if best is None or inertia < best:
                self.labels_, self.cluster_centers_ = labels.clone(), centers.clone()
                best = inertia
                self.inertia_, self.n_iter_ = inertia, n_iters


# --------------------------------------------------------------------------------------
# a7 (checkable part)  constrained variant: integer cost matrix handed to the flow solver
# --------------------------------------------------------------------------------------
def constrained_int_costs(dist_sq: torch.Tensor) -> np.ndarray:
    """``local_utils/sskm_constrained.py:116`` (``torch.sqrt(dist).cpu().numpy()``) followed by
    ``:324`` (``np.around(D * 1000, 0).astype('int32')``).  The solver itself (OR-Tools) is unpinned."""
    d = torch.sqrt(dist_sq).cpu().numpy()
    return np.around(d * 1000, 0).astype('int32')


def predict(X: torch.Tensor, centers: torch.Tensor, batch_size=None) -> torch.Tensor:
    """The reference's predict-equivalent: ``gcd/methods/clustering/k_means.py:185-186``."""
    return pairwise_distance(X, centers, batch_size).argmin(dim=-1)
