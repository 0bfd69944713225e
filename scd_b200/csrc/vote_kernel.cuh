// Per-cluster name voting and small layout helpers.
//
//   main_unsup.py:575-577 / main_ptsup.py:636-638:
//       cluster_to_counter[i] = Counter(name_idx_top5[u_preds == i, :top_k].view(-1))   (ptsup: minus known names)
//   main_unsup.py:582 / main_ptsup.py:644:   cluster_to_counter[i].most_common(num_common_vote)
//
// One CTA per cluster builds the cluster's name histogram in a shared-memory open-addressing table
// (name -> count, first flattened position) and then extracts the M most common names.  Python's
// Counter.most_common orders equal counts by first insertion, i.e. by the first position in the
// row-major flattening of the cluster's [rows, top_k] block; rows appear there in ascending row id,
// so  first = row * top_k + j  reproduces that order whatever order the rows are visited in.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace scd {

constexpr int kVoteSlots = 16384;                       // capacity of the shared-memory table
constexpr int kVoteThreads = 512;
constexpr int kVoteHistBins = 512;                      // histogram of per-name counts (last bin: >= 511)
constexpr int kVoteMaxCand = 4096;                      // candidates (count >= the M-th largest count) kept as a dense list
constexpr int kVoteSmemBytes = (3 * kVoteSlots + kVoteHistBins + kVoteMaxCand) * 4;
// bytes of the optional global-memory spill tables: a cluster with more than kVoteSlots / 2 entries (rows * k_used)
// cannot be guaranteed to fit the shared-memory table, so it builds its histogram in a private 2 * entries-slot table
// carved out of this buffer at 3 ints per slot (the cluster sizes sum to N, so 6 * N * k_used ints always suffice)
__host__ __device__ inline size_t vote_spill_bytes(long long N, int k_used) { return (size_t)(6ll * N * k_used) * sizeof(int) + 64; }

// Names are read through (pointer, row stride): int64 [N, k_total] top-k indices, or the name columns of the packed
// int32 vote records [N][1 + k] of the multi-GPU path.
// SEG = true (row-sharded vote without a second sort): `topk_idx` is the gathered record array
// [world * per][1 + k] int32 whose rank-r block holds that rank's records IN ITS OWN LABEL-SORTED ORDER,
// record = [global row id, name_0 .. name_(k-1)], and seg_offsets [world][K + 1] are the ranks' M-step offsets: the rows of
// cluster c are the `world` segments  r * per + [seg_offsets[r][c], seg_offsets[r][c + 1])  - "histograms add" by walking
// sorted runs; out_rows[c] gets the cluster's row count.
constexpr int kVoteMaxSeg = 16;

template <typename IdxT, bool SEG>
__global__ void __launch_bounds__(kVoteThreads)
vote_kernel(const IdxT* __restrict__ topk_idx, long long idx_stride, int k_used,
            const int* __restrict__ order, const int* __restrict__ offsets, int K,
            const long long* __restrict__ excluded, int n_excluded, int M,
            long long* __restrict__ out_names, int* __restrict__ out_counts, int* __restrict__ out_distinct,
            int* __restrict__ overflow_flag, int* __restrict__ spill,
            const int* __restrict__ seg_offsets, int seg_world, long long seg_per, int* __restrict__ out_rows) {
  extern __shared__ int vote_sh[];
  int* hist = vote_sh + 3 * kVoteSlots;
  unsigned* cand = reinterpret_cast<unsigned*>(vote_sh + 3 * kVoteSlots + kVoteHistBins);
  __shared__ int n_distinct, n_cand, c_star;
  __shared__ unsigned long long red[kVoteThreads / 32];
  __shared__ int red_slot[kVoteThreads / 32];
  __shared__ unsigned long long chosen_key;
  __shared__ int chosen_slot;

  const int c = blockIdx.x;
  __shared__ int seg_pre[kVoteMaxSeg + 1];          // SEG: rows of the cluster before segment r
  __shared__ long long seg_base[kVoteMaxSeg];       // SEG: record index of segment r's first row
  __shared__ int seg_p0;
  int p0, p1;
  if (SEG) {
    if (threadIdx.x == 0) {
      int run = 0, before = 0;
      for (int r = 0; r < seg_world; ++r) {
        const int a = seg_offsets[(long long)r * (K + 1) + c], b = seg_offsets[(long long)r * (K + 1) + c + 1];
        seg_pre[r] = run; seg_base[r] = (long long)r * seg_per + a;
        run += b - a; before += a;
      }
      seg_pre[seg_world] = run;
      seg_p0 = before;                               // rows of lower clusters over all ranks: this cluster's slice of the spill buffer
      if (out_rows) out_rows[c] = run;
    }
    __syncthreads();
    p0 = seg_p0; p1 = p0 + seg_pre[seg_world];
  } else {
    p0 = offsets[c]; p1 = offsets[c + 1];
  }
  const long long n_entries = (long long)(p1 - p0) * k_used;
  // where the table lives: shared memory when 2 * entries slots fit (it then can never fill up), else the spill buffer
  const bool in_smem = 2 * n_entries <= kVoteSlots || spill == nullptr;
  const unsigned nslots = in_smem ? (unsigned)min((long long)kVoteSlots, max(2 * n_entries, 64ll)) : (unsigned)(2 * n_entries);
  int* keys = in_smem ? vote_sh : spill + 6ll * p0 * k_used;
  int* cnts = keys + nslots;
  unsigned* firsts = reinterpret_cast<unsigned*>(keys + 2ll * nslots);

  const volatile int* vkeys = keys;
  const volatile unsigned* vfirsts = firsts;
  for (unsigned s = threadIdx.x; s < nslots; s += blockDim.x) { keys[s] = -1; cnts[s] = 0; firsts[s] = 0xFFFFFFFFu; }
  for (int b = threadIdx.x; b < kVoteHistBins; b += blockDim.x) hist[b] = 0;
  if (threadIdx.x == 0) { n_distinct = 0; n_cand = 0; c_star = 1; chosen_key = ~0ull; }
  __syncthreads();

  // Entries are gathered kVoteBatch at a time: the row ids of a batch are loaded first, then all its names (independent
  // loads, two round trips per batch) and only then inserted.  One entry per trip - load the row id, then its name, then
  // probe - left a thread of the 6 k-entry clusters of C2 waiting on ~25 dependent memory round trips (42 us per launch).
  constexpr int kVoteBatch = 8;
  const int n_ent = (int)n_entries;              // N * k_used < 2^31 (checked by the host)
  for (int base = 0; base < n_ent; base += kVoteBatch * (int)blockDim.x) {
    int row[kVoteBatch], jj[kVoteBatch];
    long long name64[kVoteBatch];
#pragma unroll
    for (int u = 0; u < kVoteBatch; ++u) {
      const int e = base + u * (int)blockDim.x + (int)threadIdx.x;
      row[u] = -1;
      if (e < n_ent) {
        const int r = e / k_used; jj[u] = e - r * k_used;
        if (SEG) {                                   // r-th row of the cluster -> its segment -> its record
          int sgm = 0;
          while (sgm + 1 < seg_world && r >= seg_pre[sgm + 1]) ++sgm;
          const long long pos = seg_base[sgm] + (r - seg_pre[sgm]);
          const IdxT* rp = topk_idx + pos * idx_stride;
          row[u] = (int)rp[0];                       // global row id: first positions order the ties exactly as one rank would
          name64[u] = (long long)rp[1 + jj[u]];
        } else {
          row[u] = order[p0 + r];
        }
      }
    }
    if (!SEG) {
#pragma unroll
      for (int u = 0; u < kVoteBatch; ++u)
        name64[u] = row[u] >= 0 ? (long long)topk_idx[(long long)row[u] * idx_stride + jj[u]] : -1;
    } else {
#pragma unroll
      for (int u = 0; u < kVoteBatch; ++u) if (row[u] < 0) name64[u] = -1;
    }
#pragma unroll
    for (int u = 0; u < kVoteBatch; ++u) {
      if (name64[u] < 0) continue;
      bool skip = false;
      for (int x = 0; x < n_excluded; ++x) skip |= (excluded[x] == name64[u]);
      if (skip) continue;
      const int name = (int)name64[u];
      const unsigned first = (unsigned)row[u] * (unsigned)k_used + (unsigned)jj[u];
      unsigned h = (unsigned)(((unsigned long long)((unsigned)name * 2654435761u) * nslots) >> 32);
      unsigned probes = 0;
      // Shared-memory atomics cost ~2 cycles per LANE whatever the address, so they - not the gathers - set the pace of
      // a 6 k-entry cluster.  The rows of a cluster mostly repeat names that are already in the table: a plain load finds
      // the slot (no CAS), and the first-position minimum is only attempted when it would lower the stored value.
      // One atomic per entry (the count) instead of three.
      while (true) {
        int cur = vkeys[h];
        if (cur == -1) {
          cur = atomicCAS(&keys[h], -1, name);
          if (cur == -1) { atomicAdd(&n_distinct, 1); cur = name; }
        }
        if (cur == name) {
          atomicAdd(&cnts[h], 1);
          if (first < vfirsts[h]) atomicMin(&firsts[h], first);
          break;
        }
        if (++h == nslots) h = 0;
        if (++probes >= nslots) { atomicExch(overflow_flag, 1); break; }       // only without a spill buffer
      }
    }
  }
  __syncthreads();

  // The M most common names are among the names whose count reaches the M-th largest count c*: histogram of the
  // counts -> c* -> dense candidate list (typically M .. 2 M entries), then M selection rounds by ONE warp without
  // block-wide barriers.  (Round 1 ran M block-wide arg-max rounds over every live slot: 57 us per launch.)
  for (unsigned s = threadIdx.x; s < nslots; s += blockDim.x)
    if (keys[s] >= 0) atomicAdd(&hist[min(cnts[s], kVoteHistBins - 1)], 1);
  __syncthreads();
  if (threadIdx.x < 32) {
    constexpr int kPer = kVoteHistBins / 32;
    const int lane = threadIdx.x;
    int mine = 0;
    for (int b = 0; b < kPer; ++b) mine += hist[lane * kPer + b];
    int suffix = mine;                                   // inclusive suffix sum over lanes >= this one
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_down_sync(0xffffffffu, suffix, off);
      if (lane + off < 32) suffix += o;
    }
    if (suffix >= M && suffix - mine < M) {              // the M-th largest count falls into this lane's bins
      int run = suffix - mine;
      for (int b = kPer - 1; b >= 0; --b) {
        run += hist[lane * kPer + b];
        if (run >= M) { c_star = max(lane * kPer + b, 1); break; }
      }
    }
    if (lane == 0) out_distinct[c] = n_distinct;
  }
  __syncthreads();
  const int cs = c_star;
  for (unsigned s = threadIdx.x; s < nslots; s += blockDim.x)
    if (keys[s] >= 0 && min(cnts[s], kVoteHistBins - 1) >= cs) {
      const int pos = atomicAdd(&n_cand, 1);
      if (pos < kVoteMaxCand) cand[pos] = s;
    }
  __syncthreads();
  const int nc = n_cand;
  auto key_of = [&](unsigned s) {
    return ((unsigned long long)(unsigned)cnts[s] << 32) | (unsigned long long)(0xFFFFFFFFu - firsts[s]);
  };

  if (nc <= 64) {
    // ---- one warp: M rounds of "largest (count, earliest first) strictly below the previous pick".  Only for short
    // lists: every round walks the whole list with three dependent shared-memory loads per candidate, and with the
    // long tie groups real clusters have (hundreds of names seen c* times) a single warp spent 60 k cycles here while
    // the other fifteen had exited - 80 % of the kernel (ncu source view of round 2a).  Longer lists take the
    // block-wide rounds below.
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    unsigned long long limit = ~0ull;
    for (int m = 0; m < M; ++m) {
      unsigned long long best = 0ull; int best_slot = -1;
      for (int i = lane; i < nc; i += 32) {
        const unsigned s = cand[i];
        const unsigned long long key = key_of(s);
        if (key < limit && (best_slot < 0 || key > best)) { best = key; best_slot = (int)s; }
      }
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int os = __shfl_xor_sync(0xffffffffu, best_slot, off);
        if (os >= 0 && (best_slot < 0 || ob > best)) { best = ob; best_slot = os; }
      }
      if (best_slot < 0) {                                  // fewer than M distinct names: pad
        for (int r = m + lane; r < M; r += 32) { out_names[(long long)c * M + r] = -1; out_counts[(long long)c * M + r] = 0; }
        break;
      }
      if (lane == 0) { out_names[(long long)c * M + m] = (long long)keys[best_slot]; out_counts[(long long)c * M + m] = cnts[best_slot]; }
      limit = best;
    }
    return;
  }

  // ---- many ties at c* (e.g. thousands of names seen once): block-wide rounds over the candidate list, or over the
  // whole table when even the list overflowed
  const bool use_list = nc <= kVoteMaxCand;
  const unsigned n_items = use_list ? (unsigned)nc : nslots;
  for (int m = 0; m < M; ++m) {
    const unsigned long long limit = chosen_key;
    unsigned long long best = 0ull; int best_slot = -1;
    for (unsigned i = threadIdx.x; i < n_items; i += blockDim.x) {
      const unsigned s = use_list ? cand[i] : i;
      if (!use_list && (keys[s] < 0 || min(cnts[s], kVoteHistBins - 1) < cs)) continue;
      const unsigned long long key = key_of(s);
      if (key < limit && (best_slot < 0 || key > best)) { best = key; best_slot = (int)s; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long ob = __shfl_down_sync(0xffffffffu, best, off);
      const int os = __shfl_down_sync(0xffffffffu, best_slot, off);
      if (os >= 0 && (best_slot < 0 || ob > best)) { best = ob; best_slot = os; }
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = best; red_slot[threadIdx.x >> 5] = best_slot; }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = 0ull; int bs = -1;
      for (int w = 0; w < kVoteThreads / 32; ++w)
        if (red_slot[w] >= 0 && (bs < 0 || red[w] > b)) { b = red[w]; bs = red_slot[w]; }
      chosen_slot = bs;
      if (bs >= 0) chosen_key = b;
      out_names[(long long)c * M + m] = bs >= 0 ? (long long)keys[bs] : -1;
      out_counts[(long long)c * M + m] = bs >= 0 ? cnts[bs] : 0;
    }
    __syncthreads();
    if (chosen_slot < 0) {
      for (int r = m + 1 + threadIdx.x; r < M; r += blockDim.x) { out_names[(long long)c * M + r] = -1; out_counts[(long long)c * M + r] = 0; }
      break;
    }
  }
}

// packed vote records of the multi-GPU path: rec[i] = [label_i, name_i0 .. name_i(k-1)] as int32, so ONE all-gather
// moves a rank's labels and top-k names (24 bytes per row at k = 5 instead of 48 in two collectives)
__global__ void pack_vote_records_kernel(const long long* __restrict__ labels, const long long* __restrict__ idx, int k_total,
                                         int k_used, long long n, int* __restrict__ rec) {
  const int w = 1 + k_used;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int* dst = rec + i * w;
    dst[0] = (int)labels[i];
    for (int j = 0; j < k_used; ++j) dst[1 + j] = (int)idx[i * k_total + j];
  }
}

// [D, V] (row stride ldw, fp32 or bf16) -> [V, D] bf16: the one-off re-layout of the reference's
// V-contiguous zeroshot_weights (local_utils/clip_lang_util.py:107) into the K-major B operand.
template <typename T>
__global__ void transpose_to_bf16_kernel(const T* __restrict__ W, int D, long long V, long long ldw, __nv_bfloat16* __restrict__ Wt) {
  __shared__ float tile[32][33];
  const long long v0 = (long long)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r;
    const long long v = v0 + threadIdx.x;
    tile[r][threadIdx.x] = (d < D && v < V) ? (float)W[(long long)d * ldw + v] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long v = v0 + r;
    const int d = d0 + threadIdx.x;
    if (v < V && d < D) Wt[v * D + d] = __float2bfloat16_rn(tile[threadIdx.x][r]);
  }
}

__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ in, long long n, __nv_bfloat16* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// out[r, :] = Wt[sel[r], :]   (the K voted columns of zeroshot_weights, main_unsup.py:601-602)
__global__ void gather_rows_bf16_kernel(const __nv_bfloat16* __restrict__ Wt, const long long* __restrict__ sel, int n_sel, int D,
                                        long long V, __nv_bfloat16* __restrict__ out) {
  const int r = blockIdx.x;
  const long long s = sel[r];
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[(long long)r * D + d] = (s >= 0 && s < V) ? Wt[s * D + d] : __float2bfloat16_rn(0.f);
}

}  // namespace scd
