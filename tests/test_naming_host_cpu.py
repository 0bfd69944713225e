"""CPU only: the host-side steps of the voting loop - candidate union (CPython ``list(set(...))`` order), the ``w``
matrix and the Munkres assignment through the C ABI - against the oracle (itself pinned to the reference) on the
counters of the reference-generated fixture.  The device vote that produces the counters is checked in
tests/test_gpu_naming.py."""
import os

import numpy as np
import torch

from oracle import hungarian_oracle, naming_oracle
from scd_b200 import naming


def _counters(golden_dir):
    g = np.load(os.path.join(golden_dir, 'naming_small.npz'))
    idx = torch.from_numpy(g['ptsup_idx_2500'])
    preds = g['unsup_loop_preds0']
    cluster_ids = list(set(preds))
    return g, idx, preds, cluster_ids, naming_oracle.vote(idx, preds, cluster_ids, 5)


def test_candidates_and_assignment_match_the_oracle(golden_dir):
    g, idx, preds, cluster_ids, c2c = _counters(golden_dir)
    for m in (20, 4, 1):
        assert naming.voted_candidates(c2c, cluster_ids, m) == naming_oracle.voted_candidates(c2c, cluster_ids, m)
    uniq = naming.voted_candidates(c2c, cluster_ids, 20)
    for num_common in (4, 1, 20):
        ind, w = naming.assign_name(uniq, c2c, num_common=num_common)
        ind_o, w_o = naming_oracle.assign_name(uniq, c2c, num_common=num_common)
        assert np.array_equal(w, w_o) and np.array_equal(ind, ind_o)
    # first round of the reference's loop: the voted names it recorded
    ind, _ = naming.assign_name(uniq, c2c, num_common=4)
    voted = [int(uniq[x[1]]) for x in ind[:len(cluster_ids)]]
    assert voted == g['unsup_loop_voted_0'].tolist() and len(uniq) == int(g['unsup_loop_nuniq_0'])


def test_linear_assignment_matches_oracle_on_rectangular_and_tied_costs():
    rng = np.random.RandomState(3)
    for shape in [(1, 1), (6, 6), (5, 9), (9, 5), (30, 30)]:
        for hi in (2, 40):
            cost = rng.randint(0, hi, size=shape)
            assert np.array_equal(naming.linear_assignment(cost.copy()), hungarian_oracle.linear_assignment(cost.copy()))
