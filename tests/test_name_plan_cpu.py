"""CPU only: invariants of the work partition of the fused scoring/top-k launch (scd_name_topk_plan) - whole waves of
256-row blocks sweep the vocabulary in one item, the tail wave is cut into non-empty vocabulary chunks that cover every
tile once, and the cut never costs more than not cutting.  Without a GPU the planner assumes 148 SMs (74 CTA pairs)."""
import numpy as np
import pytest

from scd_b200 import _lib

PAIRS = 74
TILE = 224


def _plan(n, v, k=5):
    out = np.zeros(6, dtype=np.int32)
    assert _lib.load().scd_name_topk_plan(n, v, k, out.ctypes.data) == 0
    return dict(zip(('row_blocks', 'tiles', 'full_rb', 'vsplit', 'tiles_per_chunk', 'pairs'), (int(x) for x in out)))


@pytest.mark.parametrize('n', [1, 255, 256, 257, 18944, 18945, 15875, 31750, 63500, 127000, 1280000])
@pytest.mark.parametrize('v', [1, 100, 224, 225, 11000, 21000, 82000, 100000])
def test_partition_invariants(n, v):
    p = _plan(n, v)
    assert p['row_blocks'] == -(-n // 256) and p['tiles'] == -(-v // TILE)
    assert p['full_rb'] % PAIRS == 0 and 0 <= p['row_blocks'] - p['full_rb'] < PAIRS
    tail = p['row_blocks'] - p['full_rb']
    s, tpc = p['vsplit'], p['tiles_per_chunk']
    assert 1 <= s <= 16 and tpc >= 1
    if tail:
        assert (s - 1) * tpc < p['tiles'] <= s * tpc              # every chunk non-empty, all tiles covered once
        waves = -(-tail * s // PAIRS)
        assert waves * tpc <= -(-tail // PAIRS) * p['tiles'] + 1e-9 or s == 1       # never worse than one chunk
    else:
        assert s == 1
    items = p['full_rb'] + tail * s
    assert p['pairs'] == min(PAIRS, max(1, items))


def test_the_tail_wave_of_the_bench_workloads_is_cut_where_it_pays():
    # C2 on one GPU: 496 row blocks = 6 whole waves + 52; at N = 8 ranks the 62 row blocks are one partial wave
    assert _plan(127000, 21000)['full_rb'] == 444 and _plan(127000, 21000)['vsplit'] > 1
    p8 = _plan(15875, 21000)
    assert p8['full_rb'] == 0 and p8['vsplit'] == 7 and p8['tiles_per_chunk'] == 14      # 6 waves x 14 tiles = 0.89 sweeps
    assert _plan(18944, 21000)['vsplit'] == 1                                             # exactly one whole wave


def test_plan_rejects_bad_arguments():
    out = np.zeros(6, dtype=np.int32)
    lib = _lib.load()
    assert lib.scd_name_topk_plan(0, 10, 5, out.ctypes.data) != 0
    assert lib.scd_name_topk_plan(10, 10, 9, out.ctypes.data) != 0
    assert b'scd_name_topk_plan' in lib.scd_last_error()
