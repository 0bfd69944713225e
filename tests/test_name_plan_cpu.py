"""CPU only: invariants of the work partition of the fused scoring/top-k launch (scd_name_topk_plan / _plan_pair): the
(256-row block, 224-name tile) space is cut into one contiguous range per CTA pair; every tile is covered exactly once,
pair loads differ by at most one tile, the pieces of a row block are numbered 0, 1, .. in pair order (= the partial-list
slots the merge reads) and never exceed the slot count the workspace is sized for.  Without a GPU the planner assumes
148 SMs (74 CTA pairs)."""
import numpy as np
import pytest

from scd_b200 import _lib

PAIRS = 74
TILE = 224


def _plan(n, v, k=5):
    out = np.zeros(6, dtype=np.int32)
    assert _lib.load().scd_name_topk_plan(n, v, k, out.ctypes.data) == 0
    return dict(zip(('row_blocks', 'tiles', 'pairs', 'slots', 'min_tiles', 'most_items'), (int(x) for x in out)))


def _items(n, v, pair, k=5):
    out = np.zeros((4096, 4), dtype=np.int32)
    cnt = _lib.load().scd_name_topk_plan_pair(n, v, k, pair, out.ctypes.data, 4096)
    assert 0 <= cnt <= 4096
    return out[:cnt]


@pytest.mark.parametrize('n', [1, 255, 256, 257, 18944, 18945, 15875, 31750, 63500, 127000, 1280000])
@pytest.mark.parametrize('v', [1, 100, 224, 225, 11000, 21000, 82000, 100000])
def test_partition_invariants(n, v):
    p = _plan(n, v)
    R, T = -(-n // 256), -(-v // TILE)
    assert p['row_blocks'] == R and p['tiles'] == T
    assert p['pairs'] == min(PAIRS, R * T)
    assert p['min_tiles'] == (R * T) // p['pairs']
    covered = np.zeros(R * T, dtype=np.int32) if R * T <= 400000 else None
    pieces_seen = {}
    total, most = 0, 0
    for q in range(p['pairs']):
        items = _items(n, v, q)
        load = int(items[:, 2].sum())
        assert p['min_tiles'] <= load <= p['min_tiles'] + 1
        total += load
        most = max(most, len(items))
        for rb, t0, nt, part in items:
            assert 0 <= rb < R and nt >= 1 and 0 <= t0 and t0 + nt <= T and 0 <= part < p['slots']
            assert part == pieces_seen.get(int(rb), 0)                 # pieces of a row block: 0, 1, 2 .. in pair order
            pieces_seen[int(rb)] = int(part) + 1
            if covered is not None:
                covered[rb * T + t0: rb * T + t0 + nt] += 1
        # a pair's range is contiguous: only its first item may start inside a row block, only its last may end inside one
        for i, (rb, t0, nt, part) in enumerate(items):
            assert t0 == 0 or i == 0
            assert t0 + nt == T or i == len(items) - 1
    assert total == R * T and most == p['most_items']
    if covered is not None:
        assert (covered == 1).all()
    assert max(pieces_seen.values()) == p['slots']


def test_small_row_shards_start_few_long_items():
    # C2 at N = 8 ranks: 63 row blocks x 94 tiles on 74 pairs = 80 tiles per pair in at most 2 items (round 1: six 14-tile items)
    p8 = _plan(15875, 21000)
    assert p8['min_tiles'] == 80 and p8['most_items'] <= 2 and p8['slots'] <= 3
    # C2 on one GPU: 497 row blocks -> 631 tiles per pair, 7-8 items, a row block is cut in at most 2 pieces
    p1 = _plan(127000, 21000)
    assert p1['min_tiles'] == 631 and p1['most_items'] <= 8 and p1['slots'] == 2
    assert _plan(18944, 21000)['slots'] == 1                              # exactly one whole row block per pair


def test_plan_rejects_bad_arguments():
    out = np.zeros(6, dtype=np.int32)
    lib = _lib.load()
    assert lib.scd_name_topk_plan(0, 10, 5, out.ctypes.data) != 0
    assert lib.scd_name_topk_plan(10, 10, 9, out.ctypes.data) != 0
    assert b'scd_name_topk_plan' in lib.scd_last_error()
    assert lib.scd_name_topk_plan_pair(1000, 1000, 5, 99, out.ctypes.data, 1) == -1
