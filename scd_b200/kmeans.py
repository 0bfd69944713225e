"""Drop-in (semi-supervised) k-means: the reference's ``K_Means`` / ``pairwise_distance`` signatures
over the sm_100a kernels in ``libscd_b200.so``.

Mirrors ``local_utils/faster_mix_k_means_pytorch.py`` (``K_Means`` :8, ``kpp`` :20, ``fit_once`` :39,
``fit_mix_once`` :77, ``fit`` :129, ``fit_mix`` :153, ``pairwise_distance`` :177) and the copy the drivers
import, ``gcd/methods/clustering/faster_mix_k_means_pytorch.py`` (adds ``mode=`` :49-59).  Same
constructor arguments, methods, result attributes and quirks (labelled rows first in ``labels_``,
``n_iter_ == len(l_targets)`` after ``fit_mix``, post-update centres paired with pre-update labels,
NaN centroid for an empty cluster).  All arithmetic runs on the GPU; inputs on the CPU are moved to
the current CUDA device and results are returned on the input's device.  No CPU fallback.

Sharded use (rows of ``X`` / ``u_feats`` block-sharded over the ranks of a ``torch.distributed``
process group, one process per GPU): pass ``process_group=`` - the M-step then all-reduces one packed
``[K*D sums | K counts | inertia]`` fp32/fp64 buffer per iteration over NCCL.
"""
from __future__ import annotations

import numpy as np
import torch
from sklearn.utils import check_random_state

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev_f32(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, on the current CUDA device (the reference feeds fp32 CUDA tensors, main_unsup.py:340)."""
    if not torch.cuda.is_available():
        raise RuntimeError('scd_b200 needs a CUDA device (there is no CPU fallback)')
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device='cuda', dtype=torch.float32).contiguous()


# ----------------------------------------------------------------------------- thin kernel wrappers
_ESTEP_WS = {}


def _estep(X, C, labels_out, inertia_acc, mindist=None, exact=False):
    lib = _lib.load()
    key = (int(C.shape[0]), int(C.shape[1]), X.device.index)
    ws = _ESTEP_WS.get(key)
    if ws is None:
        ws = _ESTEP_WS[key] = torch.empty(lib.scd_estep_workspace_bytes(key[0], key[1]), dtype=torch.uint8, device=X.device)
    _lib.check(lib.scd_estep(X.data_ptr(), X.shape[0], X.shape[1], C.data_ptr(), C.shape[0],
                             labels_out.data_ptr(), _lib.ptr(mindist), _lib.ptr(inertia_acc), int(bool(exact)),
                             ws.data_ptr(), ws.numel(), _stream()), 'scd_estep')


class _MStep:
    """Workspace + launches of the M-step for a fixed (N, D, K)."""

    def __init__(self, n, d, k, device):
        lib = _lib.load()
        self.n, self.d, self.k = n, d, k
        self.ws = torch.empty(lib.scd_mstep_workspace_bytes(n, k), dtype=torch.uint8, device=device)
        # one buffer [K*D sums | K counts | inertia] (fp32) so the row-sharded case all-reduces it in one call
        self.packed = torch.zeros(k * d + k + 1, dtype=torch.float32, device=device)
        self.sums = self.packed[:k * d].view(k, d)
        self.counts_f = self.packed[k * d:k * d + k]
        self.counts = torch.empty(k, dtype=torch.int32, device=device)
        self.norm_ws = torch.empty(max(k, 1), dtype=torch.float32, device=device)
        self.shift = torch.zeros(1, dtype=torch.float32, device=device)

    def sums_counts(self, X, labels):
        lib = _lib.load()
        _lib.check(lib.scd_mstep_sums(X.data_ptr(), labels.data_ptr(), self.n, self.d, self.k, self.sums.data_ptr(),
                                      self.counts.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream()), 'scd_mstep_sums')

    def finalize(self, c_old, c_new, counts_f=None):
        lib = _lib.load()
        _lib.check(lib.scd_finalize_centers(self.sums.data_ptr(), None if counts_f is not None else self.counts.data_ptr(),
                                            _lib.ptr(counts_f), _lib.ptr(c_old), c_new.data_ptr(),
                                            self.shift.data_ptr() if c_old is not None else None, self.k, self.d,
                                            self.norm_ws.data_ptr(), self.norm_ws.numel() * 4, _stream()), 'scd_finalize_centers')


def pairwise_distance(data1, data2, batch_size=None, *, out_device=None):
    """Reference signature (``local_utils/faster_mix_k_means_pytorch.py:177``): squared Euclidean distance
    matrix ``[N, K]``, direct form, fp32.  Like the reference, the result lives on the **CPU** when
    ``batch_size`` is given (its buffer is ``torch.zeros(N, K)``, :197) and on the input's device
    otherwise; ``out_device`` (extension) overrides that.  ``batch_size`` itself is irrelevant here: the
    kernel tiles internally and never materialises the ``B x K x D`` temporary."""
    src_device = data1.device if torch.is_tensor(data1) else torch.device('cpu')
    X, Cm = _dev_f32(data1), _dev_f32(data2)
    if X.dim() != 2 or Cm.dim() != 2 or X.shape[1] != Cm.shape[1]:
        raise ValueError(f'pairwise_distance expects [N,D] and [K,D], got {tuple(X.shape)} and {tuple(Cm.shape)}')
    out = torch.empty(X.shape[0], Cm.shape[0], dtype=torch.float32, device=X.device)
    if X.shape[0] and Cm.shape[0]:
        lib = _lib.load()
        _lib.check(lib.scd_pairwise_distance(X.data_ptr(), X.shape[0], X.shape[1], Cm.data_ptr(), Cm.shape[0],
                                             out.data_ptr(), None, _stream()), 'scd_pairwise_distance')
    target = out_device if out_device is not None else (torch.device('cpu') if batch_size is not None else src_device)
    return out.to(target)


def constrained_int_costs(data1, data2):
    """int32 ``round(1000 * sqrt(dist))`` cost matrix the size-constrained variant hands to its flow solver
    (``local_utils/sskm_constrained.py:116`` + ``:324``), produced directly by the distance kernel."""
    X, Cm = _dev_f32(data1), _dev_f32(data2)
    out = torch.empty(X.shape[0], Cm.shape[0], dtype=torch.int32, device=X.device)
    if X.shape[0] and Cm.shape[0]:
        lib = _lib.load()
        _lib.check(lib.scd_pairwise_distance(X.data_ptr(), X.shape[0], X.shape[1], Cm.data_ptr(), Cm.shape[0],
                                             None, out.data_ptr(), _stream()), 'scd_pairwise_distance')
    return out


def predict(X, cluster_centers, batch_size=None):
    """The reference has no ``K_Means.predict``; its predict-equivalent is
    ``pairwise_distance(X, centers, bs).argmin(-1)`` (``gcd/methods/clustering/k_means.py:185-186``).
    Here: the fused E-step kernel (no [N,K] matrix)."""
    src_device = X.device if torch.is_tensor(X) else torch.device('cpu')
    Xd, Cd = _dev_f32(X), _dev_f32(cluster_centers)
    labels = torch.empty(Xd.shape[0], dtype=torch.int64, device=Xd.device)
    if Xd.shape[0]:
        _estep(Xd, Cd, labels, None)
    return labels.to(src_device)


class K_Means:
    """Same constructor as the reference (``faster_mix_k_means_pytorch.py:9``; gcd copy appends ``mode``).
    ``process_group`` is the only extension (row-sharded multi-GPU)."""

    def __init__(self, k=3, tolerance=1e-4, max_iterations=100, init='k-means++', n_init=10, random_state=None,
                 n_jobs=None, pairwise_batch_size=None, mode=None, *, process_group=None):
        self.k = k
        self.tolerance = tolerance
        self.max_iterations = max_iterations
        self.init = init
        self.n_init = n_init
        self.random_state = random_state
        self.n_jobs = n_jobs                      # accepted for signature parity; restarts run sequentially on the GPU
        self.pairwise_batch_size = pairwise_batch_size
        self.mode = mode
        self.process_group = process_group

    # ------------------------------------------------------------------ helpers
    def _allreduce(self, mstep: _MStep, inertia_acc: torch.Tensor):
        """one packed all-reduce of [K*D sums | K counts | inertia] when rows are sharded."""
        if self.process_group is None:
            return None
        from . import dist as sdist
        lib = _lib.load()
        _lib.check(lib.scd_pack_counts_inertia(mstep.counts.data_ptr(), inertia_acc.data_ptr(), mstep.k,
                                               mstep.counts_f.data_ptr(), _stream()), 'scd_pack_counts_inertia')
        sdist.allreduce_packed(mstep.packed, self.process_group)
        inertia_acc.copy_(mstep.packed[-1:])
        return mstep.counts_f

    def _assign(self, X, centers, labels_out, inertia_acc):
        """E-step of one iteration: ``labels_out[i] = argmin_k ||X_i - c_k||^2`` and ``inertia_acc += sum of the
        minima``.  (The size-constrained subclass overrides this with the min-cost-flow assignment.)"""
        _estep(X, centers, labels_out, inertia_acc)

    def _lloyd(self, X_assign, X_all, labels, l_num, centers, l_feats=None):
        """The iteration loop shared by fit_once (:56-74) and fit_mix_once (:102-126).

        X_assign: rows that get (re)assigned each iteration (all of X, or u_feats); X_all: rows the M-step
        averages (X, or cat(l_feats, u_feats)); labels: int64 [len(X_all)], first l_num entries fixed."""
        dev = X_all.device
        k, d = self.k, X_all.shape[1]
        mstep = _MStep(X_all.shape[0], d, k, dev)
        inertia_acc = torch.zeros(1, dtype=torch.float64, device=dev)
        c_cur = centers.clone().contiguous()
        c_new = torch.empty_like(c_cur)
        host = torch.empty(2, dtype=torch.float64).pin_memory()
        best_labels = best_inertia = best_centers = None
        n_done = 0
        u_view = labels[l_num:]
        for it in range(self.max_iterations):
            n_done = it + 1
            inertia_acc.zero_()
            if X_assign.shape[0]:
                self._assign(X_assign, c_cur, u_view, inertia_acc)                 # :58-60 / :105-107,:111
            if l_num:
                lib = _lib.load()
                _lib.check(lib.scd_labelled_inertia(l_feats.data_ptr(), labels.data_ptr(), l_num, d, c_cur.data_ptr(), k,
                                                    inertia_acc.data_ptr(), _stream()), 'scd_labelled_inertia')   # :108-110
            mstep.sums_counts(X_all, labels)                                       # :61-64 / :113-116
            counts_f = self._allreduce(mstep, inertia_acc)
            mstep.finalize(c_cur, c_new, counts_f)                                 # divide + :71 / :123
            host[0:1].copy_(inertia_acc, non_blocking=True)
            host[1:2].copy_(mstep.shift.double(), non_blocking=True)
            torch.cuda.current_stream().synchronize()                              # the one host sync per iteration
            inertia, shift = float(host[0]), float(host[1])
            inertia32 = float(np.float32(inertia))
            if best_inertia is None or inertia32 < best_inertia:                   # :66-69 / :118-121
                best_labels, best_centers, best_inertia = labels.clone(), c_new.clone(), inertia32
            c_cur, c_new = c_new, c_cur
            if np.float32(shift) ** 2 < self.tolerance:                            # :72 / :124
                break
        return best_labels, torch.tensor(best_inertia, dtype=torch.float32, device=dev), best_centers, n_done

    # ------------------------------------------------------------------ reference API
    def kpp(self, X, pre_centers=None, k=10, random_state=None):
        """k-means++ seeding, ``faster_mix_k_means_pytorch.py:20-36`` (gcd copy :82-110 with the "no
        candidate" guard).  Keeps a running min-distance vector and only measures the newly added centre
        (``scd_kpp_update``: one N x D pass per centre instead of the reference's N x c x D); draws
        ``r = random_state.rand()`` from the same host RNG stream, one draw per centre, and resolves
        ``first index with cumsum(d2 / sum(d2)) >= r`` on the device (``scd_kpp_select``).  The picked row index
        stays on the device, so the whole seeding is a stream of launches with a single host sync at the end."""
        rs = check_random_state(random_state)
        Xd = _dev_f32(X)
        n, d = int(Xd.shape[0]), int(Xd.shape[1])
        dev = Xd.device
        if pre_centers is not None:
            Cc = _dev_f32(pre_centers).view(-1, d)
        else:
            Cc = Xd[rs.randint(0, len(Xd))].view(1, -1)                            # :25
        centers = torch.empty(max(k, Cc.shape[0]), d, dtype=torch.float32, device=dev)
        n_have = int(Cc.shape[0])
        centers[:n_have] = Cc
        if n_have >= k or n == 0:
            return centers[:n_have]
        lib = _lib.load()
        ws = torch.empty(lib.scd_kpp_workspace_bytes(n), dtype=torch.uint8, device=dev)
        d2 = torch.empty(n, dtype=torch.float32, device=dev)
        pick = torch.full((1,), -1, dtype=torch.int64, device=dev)
        no_hit = torch.zeros(1, dtype=torch.int32, device=dev)
        # distances to the centres we start from (:28-30): fused fp32 E-step, min-distance output only
        scratch_labels = torch.empty(n, dtype=torch.int64, device=dev)
        _estep(Xd, centers[:n_have].contiguous(), scratch_labels, None, mindist=d2, exact=True)
        sums_valid = 0
        while n_have < k:
            r = float(rs.rand())                                                   # :33 (one host draw per added centre)
            _lib.check(lib.scd_kpp_select(d2.data_ptr(), n, sums_valid, r, pick.data_ptr(), no_hit.data_ptr(), ws.data_ptr(),
                                          ws.numel(), _stream()), 'scd_kpp_select')                          # :31-34
            _lib.check(lib.scd_kpp_update(Xd.data_ptr(), n, d, pick.data_ptr(), 0, d2.data_ptr(), centers[n_have].data_ptr(),
                                          ws.data_ptr(), ws.numel(), _stream()), 'scd_kpp_update')           # :35 + next :28-30
            sums_valid = 1
            n_have += 1
        if int(no_hit.item()) & 2:                                                 # the one host sync of the seeding
            # gcd copy :104-107 silently reuses the previous index; with none yet, both copies fail
            raise IndexError('kpp: no cumulative probability reached the draw (reference :34)')
        return centers[:n_have]

    def fit_once(self, X, random_state):
        Xd = _dev_f32(X)
        k = self.k
        if self.init == 'k-means++':
            centers = self.kpp(Xd, k=k, random_state=random_state)                 # :44
        elif self.init == 'random':
            rs = check_random_state(self.random_state)                             # :46 (self.random_state, as the reference)
            idx = rs.choice(len(Xd), k, replace=False)
            centers = Xd[torch.as_tensor(idx, device=Xd.device)].clone()           # :47-49
        else:
            centers = Xd[:k].clone()                                               # :50-52
        labels = torch.empty(Xd.shape[0], dtype=torch.int64, device=Xd.device)
        return self._lloyd(Xd, Xd, labels, 0, centers)

    def fit_mix_once(self, u_feats, l_feats, l_targets, random_state):
        U, L = _dev_f32(u_feats), _dev_f32(l_feats)
        tg = l_targets if torch.is_tensor(l_targets) else torch.as_tensor(l_targets)
        tg = tg.to(U.device)
        l_classes, remapped = torch.unique(tg, return_inverse=True)                # :80, :93-95 (sorted unique -> 0..C_l-1)
        l_num, n_lab_classes = int(tg.numel()), int(l_classes.numel())
        cat_feats = torch.cat((L, U)).contiguous()                                 # :83
        labels = torch.full((cat_feats.shape[0],), -1, dtype=torch.int64, device=U.device)     # :88
        labels[:l_num] = remapped.view(-1).long()
        # labelled class means (:78-82) with the M-step kernels
        ms = _MStep(l_num, L.shape[1], n_lab_classes, U.device)
        ms.sums_counts(L, labels[:l_num].contiguous())
        l_centers = torch.empty(n_lab_classes, L.shape[1], dtype=torch.float32, device=U.device)
        ms.finalize(None, l_centers)
        centers = self.kpp(U, l_centers, k=self.k, random_state=random_state)      # :98
        best_labels, best_inertia, best_centers, _ = self._lloyd(U, cat_feats, labels, l_num, centers, l_feats=L)
        # :127 returns the stale loop variable of :94 + 1, i.e. the number of labelled rows
        return best_labels, best_inertia, best_centers, l_num

    def _best_of(self, run, src_device):
        rs = check_random_state(self.random_state)
        best = None
        for _ in range(self.n_init):                                               # :133-140 / :157-164
            labels, inertia, centers, n_iters = run(rs)
            if best is None or float(inertia) < best:
                self.labels_ = labels.clone().to(src_device)
                self.cluster_centers_ = centers.clone().to(src_device)
                best = float(inertia)
                self.inertia_ = inertia.to(src_device)
                self.n_iter_ = n_iters

    def fit(self, X):
        src = X.device if torch.is_tensor(X) else torch.device('cpu')
        Xd = _dev_f32(X)
        self._best_of(lambda rs: self.fit_once(Xd, rs), src)

    def fit_mix(self, u_feats, l_feats, l_targets):
        src = u_feats.device if torch.is_tensor(u_feats) else torch.device('cpu')
        U, L = _dev_f32(u_feats), _dev_f32(l_feats)
        self._best_of(lambda rs: self.fit_mix_once(U, L, l_targets, rs), src)
