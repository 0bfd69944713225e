"""CPU only: the on-disk contracts (feature / cluster-result `.pt` dicts) are byte-compatible with the files the
reference drivers write - fixtures produced by the reference's own extract_feature (oracle/gen_golden.py: gen_eval)."""
import os

import numpy as np
import pytest
import torch

from scd_b200 import features


def _same_dict(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        if a[k] is None or b[k] is None:
            assert a[k] is None and b[k] is None
            continue
        assert isinstance(a[k], np.ndarray) and isinstance(b[k], np.ndarray), k
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


def test_feature_file_written_by_the_reference_loads_and_round_trips(golden_dir, tmp_path):
    ref = torch.load(os.path.join(golden_dir, 'features_all.pt'), weights_only=False)
    d = features.load_features(os.path.join(golden_dir, 'features_all.pt'))
    _same_dict(d, ref)
    assert d['all_feats'].dtype == np.float32 and d['targets'].dtype == np.float64 and d['mask_lab'].dtype == bool
    raw = np.load(os.path.join(golden_dir, 'features_raw.npz'))
    feats = torch.nn.functional.normalize(torch.from_numpy(raw['raw']), dim=-1)            # main_unsup.py:130
    mine = features.save_features(tmp_path / 'sub' / 'mine.pt', feats, raw['lab_mask'], raw['labels'] < 4, raw['labels'])
    _same_dict(mine, ref)
    _same_dict(torch.load(tmp_path / 'sub' / 'mine.pt', weights_only=False), ref)           # what the reference would read back


def test_feature_file_errors(golden_dir, tmp_path):
    torch.save({'all_feats': np.zeros((3, 2), np.float32)}, tmp_path / 'bad.pt')
    with pytest.raises(KeyError):
        features.load_features(tmp_path / 'bad.pt')
    with pytest.raises(ValueError):
        features.save_features(tmp_path / 'bad2.pt', np.zeros((3, 2)), np.zeros(2), np.zeros(3), np.zeros(3))


def test_cluster_result_file_matches_the_reference_layout(golden_dir, tmp_path):
    ref = torch.load(os.path.join(golden_dir, 'cluster_result.pt'), weights_only=False)
    all_preds, u_preds, u_targets, mask = features.load_cluster_result(os.path.join(golden_dir, 'cluster_result.pt'))
    assert np.array_equal(all_preds, ref['all_preds']) and np.array_equal(u_preds, ref['u_preds'])
    assert u_targets.dtype == np.float64 and mask.dtype == bool
    mine = features.save_cluster_result(tmp_path / 'c.pt', torch.from_numpy(all_preds), u_preds, u_targets, mask)
    _same_dict(mine, ref)
    _same_dict(torch.load(tmp_path / 'c.pt', weights_only=False), ref)
    km = features.save_cluster_result(tmp_path / 'km.pt', None, u_preds, u_targets, mask)    # --cluster KM: all_preds stays None (:336)
    assert km['all_preds'] is None and features.load_cluster_result(tmp_path / 'km.pt')[0] is None
