"""Oracle: Kuhn-Munkres assignment with the reference's tie-breaking.

TEST INFRASTRUCTURE - not shipped, never imported by ``scd_b200``.

``local_utils/clip_lang_util.py:178`` solves ``linear_assignment(w.max() - w)`` with the pure-NumPy
Munkres that scikit-learn shipped until 0.22, vendored at ``gcd/project_utils/cluster_utils.py:234-492``.
Optimal assignments are not unique under tied integer costs, so voted names are only reproducible if
every tie is broken the same way.  This restates the six published Munkres steps with exactly the
choices the vendored code makes:
  * step 1/2 (:350-366): row-reduce, then star zeros greedily in row-major order;
  * step 3 (:369-380):   cover starred columns; finished when #stars == #rows;
  * step 4 (:383-421):   always take the *first uncovered zero in row-major order* (``argmax`` of a
                         0/1 mask), prime it; if its row holds a star (first one, lowest column)
                         cover the row / uncover that column, else augment from it;
  * step 5 (:424-473):   alternating path through the first star in the column (lowest row) and the
                         prime in that row (lowest column); flip, clear covers and primes;
  * step 6 (:476-489):   smallest uncovered value added to covered rows, subtracted from uncovered columns.
Pinned by ``tests/golden/hungarian_*.npz`` produced with the vendored reference implementation.
"""
from __future__ import annotations

import numpy as np

STAR, PRIME = 1, 2


class _Munkres:
    def __init__(self, cost):
        cost = np.atleast_2d(cost)
        self.flipped = cost.shape[1] < cost.shape[0]          # ref :300-305 work on the wide orientation
        self.c = (cost.T if self.flipped else cost).copy()
        n, m = self.c.shape
        self.row_free = np.ones(n, dtype=bool)
        self.col_free = np.ones(m, dtype=bool)
        self.mark = np.zeros((n, m), dtype=int)

    def reduce_and_star(self):
        self.c -= self.c.min(axis=1)[:, None]
        rows, cols = np.nonzero(self.c == 0)                  # row-major order
        for r, q in zip(rows, cols):
            if self.row_free[r] and self.col_free[q]:
                self.mark[r, q] = STAR
                self.row_free[r] = False
                self.col_free[q] = False
        self.row_free[:] = True
        self.col_free[:] = True

    def cover_stars(self) -> bool:
        starred = self.mark == STAR
        self.col_free[starred.any(axis=0)] = False
        return starred.sum() >= self.c.shape[0]

    def prime_until_augment(self):
        """Returns the (row, col) of a primed zero whose row has no star, or None when no
        uncovered zero is left (-> step 6)."""
        n, m = self.c.shape
        zero = (self.c == 0).astype(int)
        open_zero = zero * self.row_free[:, None].astype(int) * self.col_free[None, :].astype(int)
        while True:
            flat = int(np.argmax(open_zero))
            r, q = divmod(flat, m)
            if open_zero[r, q] == 0:
                return None
            self.mark[r, q] = PRIME
            s = int(np.argmax(self.mark[r] == STAR))
            if self.mark[r, s] != STAR:
                return r, q
            self.row_free[r] = False
            self.col_free[s] = True
            open_zero[:, s] = zero[:, s] * self.row_free.astype(int)
            open_zero[r] = 0

    def augment(self, r0, q0):
        path = [(r0, q0)]
        while True:
            q = path[-1][1]
            r = int(np.argmax(self.mark[:, q] == STAR))
            if self.mark[r, q] != STAR:
                break
            path.append((r, q))
            q2 = int(np.argmax(self.mark[r] == PRIME))
            if self.mark[r, q2] != PRIME:
                q2 = -1
            path.append((r, q2))
        for r, q in path:
            self.mark[r, q] = 0 if self.mark[r, q] == STAR else STAR
        self.row_free[:] = True
        self.col_free[:] = True
        self.mark[self.mark == PRIME] = 0

    def shift_by_min(self):
        if self.row_free.any() and self.col_free.any():
            delta = self.c[self.row_free][:, self.col_free].min()
            self.c[~self.row_free] += delta
            self.c[:, self.col_free] -= delta

    def solve(self):
        if 0 in self.c.shape:
            return np.zeros((0, 2), dtype=int)
        self.reduce_and_star()
        while not self.cover_stars():
            while True:
                hit = self.prime_until_augment()
                if hit is not None:
                    break
                self.shift_by_min()
            self.augment(*hit)
        pairs = np.array(np.nonzero(self.mark == STAR)).T
        if self.flipped:
            pairs = pairs[:, ::-1]
        return pairs


def linear_assignment(cost) -> np.ndarray:
    """ref ``cluster_utils.py:234-275``: (row, col) pairs sorted lexicographically, int dtype, shape [-1, 2]."""
    pairs = _Munkres(cost).solve().tolist()
    pairs.sort()
    out = np.array(pairs, dtype=int)
    out.shape = (-1, 2)
    return out
