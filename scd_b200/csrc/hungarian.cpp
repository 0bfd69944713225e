// Kuhn-Munkres assignment on the host with the tie-breaking of the Munkres implementation the
// reference calls (gcd/project_utils/cluster_utils.py:234-492, reached from
// local_utils/clip_lang_util.py:178 as linear_assignment(w.max() - w)).  Optimal assignments are not
// unique under tied integer costs and the voted names depend on which optimum is returned, so every
// choice below follows that implementation's order:
//   greedy starring of zeros in row-major order; always the first uncovered zero in row-major
//   order; first star in a row = lowest column; first star in a column = lowest row.
// The reference runs this step on the CPU as well (pure NumPy); it is outside the timed naming round.
#include "../../include/scd_b200.h"

#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Munkres {
  int n, m;                       // n <= m
  std::vector<int64_t> c;
  std::vector<uint8_t> row_free, col_free, mark;   // mark: 0 none, 1 star, 2 prime
  int64_t& C(int i, int j) { return c[(size_t)i * m + j]; }
  uint8_t& M(int i, int j) { return mark[(size_t)i * m + j]; }

  void reduce_and_star() {
    for (int i = 0; i < n; ++i) {
      int64_t mn = std::numeric_limits<int64_t>::max();
      for (int j = 0; j < m; ++j) mn = C(i, j) < mn ? C(i, j) : mn;
      for (int j = 0; j < m; ++j) C(i, j) -= mn;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
        if (C(i, j) == 0 && row_free[i] && col_free[j]) { M(i, j) = 1; row_free[i] = 0; col_free[j] = 0; }
    std::fill(row_free.begin(), row_free.end(), 1);
    std::fill(col_free.begin(), col_free.end(), 1);
  }
  bool cover_stars() {
    int stars = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
        if (M(i, j) == 1) { col_free[j] = 0; ++stars; }
    return stars >= n;
  }
  // returns true with (r, q) = primed zero whose row has no star; false when no uncovered zero is left
  bool prime_until_augment(int* r_out, int* q_out) {
    while (true) {
      int r = -1, q = -1;
      for (int i = 0; i < n && r < 0; ++i) {
        if (!row_free[i]) continue;
        for (int j = 0; j < m; ++j)
          if (col_free[j] && C(i, j) == 0) { r = i; q = j; break; }
      }
      if (r < 0) return false;
      M(r, q) = 2;
      int s = -1;
      for (int j = 0; j < m; ++j) if (M(r, j) == 1) { s = j; break; }
      if (s < 0) { *r_out = r; *q_out = q; return true; }
      row_free[r] = 0;
      col_free[s] = 1;
    }
  }
  void augment(int r0, int q0) {
    std::vector<std::pair<int, int>> path;
    path.emplace_back(r0, q0);
    while (true) {
      const int q = path.back().second;
      int r = -1;
      for (int i = 0; i < n; ++i) if (M(i, q) == 1) { r = i; break; }
      if (r < 0) break;
      path.emplace_back(r, q);
      int q2 = -1;
      for (int j = 0; j < m; ++j) if (M(r, j) == 2) { q2 = j; break; }
      path.emplace_back(r, q2);
    }
    for (auto& pq : path) M(pq.first, pq.second) = (M(pq.first, pq.second) == 1) ? 0 : 1;
    std::fill(row_free.begin(), row_free.end(), 1);
    std::fill(col_free.begin(), col_free.end(), 1);
    for (auto& v : mark) if (v == 2) v = 0;
  }
  void shift_by_min() {
    int64_t delta = std::numeric_limits<int64_t>::max();
    bool any = false;
    for (int i = 0; i < n; ++i) {
      if (!row_free[i]) continue;
      for (int j = 0; j < m; ++j)
        if (col_free[j]) { any = true; if (C(i, j) < delta) delta = C(i, j); }
    }
    if (!any) return;
    for (int i = 0; i < n; ++i)
      if (!row_free[i]) for (int j = 0; j < m; ++j) C(i, j) += delta;
    for (int j = 0; j < m; ++j)
      if (col_free[j]) for (int i = 0; i < n; ++i) C(i, j) -= delta;
  }
};

}  // namespace

extern "C" __attribute__((visibility("default"))) int scd_linear_assignment(const int64_t* cost, int n_rows, int n_cols, int64_t* out_pairs, int* n_pairs) {
  if (n_rows < 0 || n_cols < 0 || !n_pairs) return 1;
  *n_pairs = 0;
  if (n_rows == 0 || n_cols == 0) return 0;
  if (!cost || !out_pairs) return 1;
  const bool flipped = n_cols < n_rows;          // work on the wide orientation, swap back at the end
  Munkres mk;
  mk.n = flipped ? n_cols : n_rows;
  mk.m = flipped ? n_rows : n_cols;
  mk.c.resize((size_t)mk.n * mk.m);
  for (int i = 0; i < n_rows; ++i)
    for (int j = 0; j < n_cols; ++j) {
      const int64_t v = cost[(size_t)i * n_cols + j];
      if (flipped) mk.c[(size_t)j * mk.m + i] = v; else mk.c[(size_t)i * mk.m + j] = v;
    }
  mk.row_free.assign(mk.n, 1);
  mk.col_free.assign(mk.m, 1);
  mk.mark.assign((size_t)mk.n * mk.m, 0);
  mk.reduce_and_star();
  while (!mk.cover_stars()) {
    int r, q;
    while (!mk.prime_until_augment(&r, &q)) mk.shift_by_min();
    mk.augment(r, q);
  }
  // pairs in (row, col) lexicographic order of the ORIGINAL orientation
  int cnt = 0;
  if (!flipped) {
    for (int i = 0; i < mk.n; ++i)
      for (int j = 0; j < mk.m; ++j)
        if (mk.M(i, j) == 1) { out_pairs[2 * cnt] = i; out_pairs[2 * cnt + 1] = j; ++cnt; }
  } else {
    for (int j = 0; j < mk.m; ++j)          // original row = j, original col = i
      for (int i = 0; i < mk.n; ++i)
        if (mk.M(i, j) == 1) { out_pairs[2 * cnt] = j; out_pairs[2 * cnt + 1] = i; ++cnt; }
  }
  *n_pairs = cnt;
  return 0;
}
