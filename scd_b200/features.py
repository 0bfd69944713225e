"""On-disk contracts either side of the naming round: the ``.pt`` files the reference drivers
exchange between feature extraction, clustering and naming - same dict keys, dtypes and shapes, so files written by
the reference load here and vice versa.

  * features     ``main_unsup.py:294-300`` / ``:305-310``: ``torch.save(extract_feature(...))`` with the dict of
                 ``extract_feature`` ``:113-146`` - ``all_feats`` float32 ``[N, D]`` (L2-normalised, ``:130``),
                 ``mask_lab`` bool ``[N]``, ``mask_cls`` bool ``[N]``, ``targets`` float64 ``[N]`` (``np.append``, ``:132``)
  * split        ``main_unsup.py:321-331``: labelled / unlabelled rows and the old-class mask of the unlabelled rows
  * cluster file ``main_unsup.py:366-374``: ``all_preds`` (None for ``--cluster KM``), ``u_preds``, ``u_targets``, ``mask``
  * vocabulary   ``main_unsup.py:389-395``: ``zeroshot_weights [D, V]`` tensor saved by ``clip_lang_util.py:107``

The reference keeps everything as host NumPy and uploads per 1024-row batch inside its loops
(``torch.from_numpy(...).cuda()``, ``main_unsup.py:522``); ``FeatureSet`` uploads once and hands out device-resident
views in the layouts the kernels want (fp32 for k-means, bf16 for the scoring kernel).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import naming

FEATURE_KEYS = ('all_feats', 'mask_lab', 'mask_cls', 'targets')
CLUSTER_KEYS = ('all_preds', 'u_preds', 'u_targets', 'mask')


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def save_features(path, all_feats, mask_lab, mask_cls, targets) -> dict:
    """Write the dict of ``extract_feature`` (``main_unsup.py:140-146``) exactly as ``torch.save(data_dict, save_dir)``
    does (``:298``).  Features are stored as given (the reference normalises before collecting them, ``:130``)."""
    data_dict = {
        'all_feats': np.ascontiguousarray(_np(all_feats), dtype=np.float32),
        'mask_lab': _np(mask_lab).astype(bool),
        'mask_cls': _np(mask_cls).astype(bool),
        'targets': _np(targets).astype(np.float64),
    }
    n = data_dict['all_feats'].shape[0]
    for k in FEATURE_KEYS[1:]:
        if data_dict[k].shape != (n,):
            raise ValueError(f'{k} has shape {data_dict[k].shape}, expected ({n},)')
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(data_dict, path)
    return data_dict


def load_features(path) -> dict:
    """``torch.load(save_dir)`` (``main_unsup.py:300``) with the keys / shapes checked."""
    data_dict = torch.load(path, weights_only=False)
    missing = [k for k in FEATURE_KEYS if k not in data_dict]
    if missing:
        raise KeyError(missing[0])                                     # what ``data_dict['...']`` would raise at :321
    n = _np(data_dict['all_feats']).shape[0]
    for k in FEATURE_KEYS[1:]:
        if _np(data_dict[k]).shape[0] != n:
            raise ValueError(f'{k} has {_np(data_dict[k]).shape[0]} entries for {n} feature rows')
    return data_dict


class FeatureSet:
    """One feature file on the device, split as ``main_unsup.py:321-331`` splits it.

    Attributes (device tensors unless noted): ``l_feats`` / ``u_feats`` fp32, ``l_targets`` / ``u_targets`` float64 (the
    dtype the drivers hand to ``fit_mix``), ``mask`` = ``mask_cls[~mask_lab]`` bool (host ndarray, as ``:329-331``),
    ``mask_lab`` (host ndarray).  ``bf16()`` gives the unlabelled rows in the scoring kernel's operand layout."""

    def __init__(self, data_dict):
        naming._require_cuda()
        all_feats = torch.from_numpy(np.ascontiguousarray(_np(data_dict['all_feats']), dtype=np.float32)).to('cuda')
        self.mask_lab = _np(data_dict['mask_lab']).astype(bool)
        mask_cls = _np(data_dict['mask_cls'])
        targets = torch.from_numpy(_np(data_dict['targets']).astype(np.float64)).to('cuda')
        lab = torch.from_numpy(self.mask_lab).to('cuda')
        self.all_feats = all_feats
        self.l_feats = all_feats[lab].contiguous()                      # :323
        self.u_feats = all_feats[~lab].contiguous()                     # :324
        self.l_targets = targets[lab].contiguous()                      # :325
        self.u_targets = targets[~lab].contiguous()                     # :326
        self.mask = mask_cls[~self.mask_lab].astype(bool)               # :329-331
        self._u_bf16 = None

    @classmethod
    def load(cls, path):
        return cls(load_features(path))

    def bf16(self) -> torch.Tensor:
        if self._u_bf16 is None:
            self._u_bf16 = naming._feats_bf16(self.u_feats)
        return self._u_bf16


def save_cluster_result(path, all_preds, u_preds, u_targets, mask) -> dict:
    """``main_unsup.py:366-371``: ``{'all_preds', 'u_preds', 'u_targets', 'mask'}`` as host NumPy arrays."""
    cluster_result = {
        'all_preds': None if all_preds is None else _np(all_preds),
        'u_preds': _np(u_preds),
        'u_targets': _np(u_targets),
        'mask': _np(mask).astype(bool),
    }
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(cluster_result, path)
    return cluster_result


def load_cluster_result(path):
    """``main_unsup.py:373-374``: returns ``(all_preds, u_preds, u_targets, mask)``."""
    cluster_result = torch.load(path, weights_only=False)
    return tuple(cluster_result[k] for k in CLUSTER_KEYS)


def load_vocabulary(path, col_offset: int = 0) -> 'naming.Vocabulary':
    """``zeroshot_weights = torch.load(...)`` (``main_unsup.py:389-395``): the ``[D, V]`` tensor
    ``zeroshot_classifier`` stacks (``clip_lang_util.py:107``), re-laid out once for the scoring kernel."""
    w = torch.load(path, weights_only=False, map_location='cpu')
    if not torch.is_tensor(w) or w.dim() != 2:
        raise ValueError('expected a [D, V] tensor of zeroshot weights')
    return naming.Vocabulary(w, col_offset=col_offset)
