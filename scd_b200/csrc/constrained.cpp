// Size-constrained assignment: the min-cost-flow step of the constrained k-means E-step,
//   local_utils/sskm_constrained.py:226-274 _labels_constrained
//     -> :277-328 minimum_cost_flow_problem_graph  (X -> C' arcs cap 1 cost round(1000*sqrt(d)); C' -> C cap size_max;
//                                                   C demands size_min; artificial node takes the rest)
//     -> :331-356 solve_min_cost_flow_graph        (OR-Tools SimpleMinCostFlow; labels = flow[i, :].argmax())
// i.e.  minimise sum_i cost[i, label_i]  subject to  size_min <= |{i : label_i = k}| <= size_max for every k.
//
// The reference hands the explicit N*K-arc graph to OR-Tools' cost-scaling push-relabel (not in the tree, not
// installed).  This solver uses the structure instead: it starts from the unconstrained optimum (row argmin) and
// repairs the size violations by successive shortest augmenting paths on the (K+1)-node residual graph
//     cluster a -> cluster b   weight min_{i in a} (cost[i,b] - cost[i,a])       (move the best item of a into b)
//     cluster k -> T / T -> k  weight 0 while the optional part of k's outflow (0 .. size_max-size_min) has room / is used
// with node potentials (dense Dijkstra, O(K^2) per unit of violation).  Every augmentation keeps reduced-cost
// optimality, so the result is an exact optimum; when no cluster violates its bounds nothing is built at all.
// Optimal label vectors are not unique under tied integer costs: parity with the reference is gated on the optimal
// total cost and on feasibility ("parity unpinned" for the solver), not on labels.
// Host code: the reference runs this step on the CPU too (the N x K cost matrix is copied to the host, :116).
#include "../../include/scd_b200.h"

#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

namespace {

struct Entry {
  int32_t key;    // cost[item, b] - cost[item, a]
  int32_t item;
};
struct EntryGreater {
  bool operator()(const Entry& x, const Entry& y) const { return x.key != y.key ? x.key > y.key : x.item > y.item; }
};

constexpr int64_t kInf = std::numeric_limits<int64_t>::max() / 4;

struct Solver {
  const int32_t* cost;
  int64_t n;
  int k;
  int64_t lo, hi;
  std::vector<int32_t> assign;
  std::vector<int64_t> x;          // items per cluster
  std::vector<int64_t> f;          // optional outflow of cluster k (0 .. hi - lo)
  std::vector<std::vector<Entry>> heap;     // [a * k + b], min-heap on (key, item); entries of items that left a are stale
  std::vector<int64_t> pi;         // potentials of the k clusters + T (index k)
  int64_t augmentations = 0;

  int32_t c(int64_t i, int j) const { return cost[i * k + j]; }

  void push_item(int32_t i, int a) {
    const int32_t base = c(i, a);
    for (int b = 0; b < k; ++b) {
      if (b == a) continue;
      auto& h = heap[(size_t)a * k + b];
      h.push_back(Entry{(int32_t)(c(i, b) - base), i});
      std::push_heap(h.begin(), h.end(), EntryGreater());
    }
  }

  // cheapest live move a -> b (kInf if a has no items)
  int64_t top(int a, int b) {
    auto& h = heap[(size_t)a * k + b];
    while (!h.empty() && assign[h.front().item] != a) {
      std::pop_heap(h.begin(), h.end(), EntryGreater());
      h.pop_back();
    }
    return h.empty() ? kInf : (int64_t)h.front().key;
  }

  int64_t excess(int v) const {      // v < k: cluster; v == k: T
    if (v < k) return x[v] - lo - f[v];
    int64_t out = 0;
    for (int j = 0; j < k; ++j) out += lo + f[j];
    return out - n;
  }

  void build() {
    heap.assign((size_t)k * k, {});
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b)
        if (a != b) heap[(size_t)a * k + b].reserve((size_t)x[a]);
    for (int64_t i = 0; i < n; ++i) {
      const int a = assign[i];
      const int32_t base = c(i, a);
      for (int b = 0; b < k; ++b)
        if (b != a) heap[(size_t)a * k + b].push_back(Entry{(int32_t)(c(i, b) - base), (int32_t)i});
    }
    for (auto& h : heap) std::make_heap(h.begin(), h.end(), EntryGreater());
  }

  // one unit from excess node s to the nearest deficit node; false if none is reachable
  bool augment(int s) {
    const int nodes = k + 1, T = k;
    std::vector<int64_t> d(nodes, kInf);
    std::vector<int> prev(nodes, -1);
    std::vector<uint8_t> done(nodes, 0);
    d[s] = 0;
    int t = -1;
    for (;;) {
      int u = -1;
      int64_t du = kInf;
      for (int v = 0; v < nodes; ++v)
        if (!done[v] && d[v] < du) { du = d[v]; u = v; }
      if (u < 0) break;
      done[u] = 1;
      if (u != s && excess(u) < 0) { t = u; break; }
      if (u == T) {
        for (int j = 0; j < k; ++j) {
          if (done[j] || f[j] <= 0) continue;                    // T -> j: undo one unit of j's optional outflow
          const int64_t nd = du + pi[T] - pi[j];
          if (nd < d[j]) { d[j] = nd; prev[j] = T; }
        }
      } else {
        if (!done[T] && f[u] < hi - lo) {                        // u -> T: one more unit of optional outflow
          const int64_t nd = du + pi[u] - pi[T];
          if (nd < d[T]) { d[T] = nd; prev[T] = u; }
        }
        if (x[u] > 0) {
          for (int b = 0; b < k; ++b) {
            if (b == u || done[b]) continue;
            const int64_t w = top(u, b);
            if (w >= kInf) continue;
            const int64_t nd = du + w + pi[u] - pi[b];
            if (nd < d[b]) { d[b] = nd; prev[b] = u; }
          }
        }
      }
    }
    if (t < 0) return false;
    const int64_t dt = d[t];
    for (int v = 0; v < nodes; ++v) pi[v] += (done[v] && d[v] < dt) ? d[v] : dt;
    // walk the path backwards and apply it
    for (int v = t; v != s;) {
      const int u = prev[v];
      if (u == T) {
        f[v] -= 1;
      } else if (v == T) {
        f[u] += 1;
      } else {
        top(u, v);                                               // make sure the front is live
        auto& h = heap[(size_t)u * k + v];
        const int32_t item = h.front().item;
        std::pop_heap(h.begin(), h.end(), EntryGreater());
        h.pop_back();
        assign[item] = v;
        x[u] -= 1;
        x[v] += 1;
        push_item(item, v);
      }
      v = u;
    }
    ++augmentations;
    return true;
  }

  // 0 ok, 2 infeasible
  int solve() {
    assign.resize((size_t)n);
    x.assign(k, 0);
    for (int64_t i = 0; i < n; ++i) {
      int best = 0;
      int32_t bv = c(i, 0);
      for (int j = 1; j < k; ++j)
        if (c(i, j) < bv) { bv = c(i, j); best = j; }
      assign[i] = best;
      x[best] += 1;
    }
    bool ok = true;
    for (int j = 0; j < k; ++j) ok = ok && x[j] >= lo && x[j] <= hi;
    if (ok) return 0;
    f.assign(k, 0);
    for (int j = 0; j < k; ++j) f[j] = std::min(std::max<int64_t>(x[j] - lo, 0), hi - lo);
    pi.assign(k + 1, 0);
    build();
    // overfull clusters first, then T (it holds the units the underfull clusters are owed)
    for (int s = 0; s <= k; ++s)
      while (excess(s) > 0)
        if (!augment(s)) return 2;
    for (int j = 0; j < k; ++j)
      if (x[j] < lo || x[j] > hi) return 2;
    return 0;
  }
};

}  // namespace

extern "C" int scd_constrained_assign(const int32_t* cost, int64_t N, int K, int64_t size_min, int64_t size_max,
                                      int32_t* labels, int64_t* total_cost, int64_t* n_augment) {
  if (N < 0 || K <= 0 || !labels || (N > 0 && !cost)) return 1;
  if (N >= (1ll << 31)) return 1;
  if (size_min < 0) size_min = 0;
  // the flow problem of sskm_constrained.py:277-328 is feasible iff K*size_min <= N <= K*size_max
  if (size_max < size_min || (int64_t)K * size_min > N || (size_max < N && (int64_t)K * size_max < N)) return 2;
  if (size_max > N) size_max = N;
  Solver s;
  s.cost = cost;
  s.n = N;
  s.k = K;
  s.lo = size_min;
  s.hi = size_max;
  const int rc = s.solve();
  if (rc != 0) return rc;
  int64_t tot = 0;
  for (int64_t i = 0; i < N; ++i) {
    labels[i] = s.assign[i];
    tot += cost[i * K + s.assign[i]];
  }
  if (total_cost) *total_cost = tot;
  if (n_augment) *n_augment = s.augmentations;
  return 0;
}
