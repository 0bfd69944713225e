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

constexpr int kVoteSlots = 16384;                       // power of two
constexpr int kVoteThreads = 512;
constexpr int kVoteSmemBytes = kVoteSlots * 14;         // keys + counts + first positions + live-slot list (u16)

__global__ void __launch_bounds__(kVoteThreads)
vote_kernel(const long long* __restrict__ topk_idx, int k_total, int k_used,
            const int* __restrict__ order, const int* __restrict__ offsets, int K,
            const long long* __restrict__ excluded, int n_excluded, int M,
            long long* __restrict__ out_names, int* __restrict__ out_counts, int* __restrict__ out_distinct,
            int* __restrict__ overflow_flag) {
  extern __shared__ int vote_sh[];
  int* keys = vote_sh;
  int* cnts = vote_sh + kVoteSlots;
  unsigned* firsts = reinterpret_cast<unsigned*>(vote_sh + 2 * kVoteSlots);
  unsigned short* live = reinterpret_cast<unsigned short*>(vote_sh + 3 * kVoteSlots);   // slots in use, dense
  __shared__ int n_live;
  __shared__ unsigned long long red[kVoteThreads / 32];
  __shared__ int red_slot[kVoteThreads / 32];
  __shared__ unsigned long long chosen_key;
  __shared__ int chosen_slot;
  __shared__ int n_distinct;

  const int c = blockIdx.x;
  for (int s = threadIdx.x; s < kVoteSlots; s += blockDim.x) { keys[s] = -1; cnts[s] = 0; firsts[s] = 0xFFFFFFFFu; }
  if (threadIdx.x == 0) { n_distinct = 0; n_live = 0; }
  __syncthreads();

  const int p0 = offsets[c], p1 = offsets[c + 1];
  const long long n_entries = (long long)(p1 - p0) * k_used;
  for (long long e = threadIdx.x; e < n_entries; e += blockDim.x) {
    const int row = order[p0 + (int)(e / k_used)];
    const int j = (int)(e % k_used);
    const long long name64 = topk_idx[(long long)row * k_total + j];
    if (name64 < 0) continue;
    bool skip = false;
    for (int x = 0; x < n_excluded; ++x) skip |= (excluded[x] == name64);
    if (skip) continue;
    const int name = (int)name64;
    const unsigned first = (unsigned)row * (unsigned)k_used + (unsigned)j;
    unsigned h = ((unsigned)name * 2654435761u) & (kVoteSlots - 1);
    int probes = 0;
    while (true) {
      const int prev = atomicCAS(&keys[h], -1, name);
      if (prev == -1) atomicAdd(&n_distinct, 1);
      if (prev == -1 || prev == name) { atomicAdd(&cnts[h], 1); atomicMin(&firsts[h], first); break; }
      h = (h + 1) & (kVoteSlots - 1);
      if (++probes >= kVoteSlots) { atomicExch(overflow_flag, 1); break; }
    }
  }
  __syncthreads();
  // dense list of the slots in use: the M selection rounds below scan n_distinct entries, not 16384 slots
  for (int s = threadIdx.x; s < kVoteSlots; s += blockDim.x)
    if (keys[s] >= 0) live[atomicAdd(&n_live, 1)] = (unsigned short)s;
  if (threadIdx.x == 0) { out_distinct[c] = n_distinct; chosen_key = ~0ull; }
  __syncthreads();
  const int nl = n_live;

  // M rounds of "largest (count, earliest first) strictly below the previous pick"
  for (int m = 0; m < M; ++m) {
    const unsigned long long limit = chosen_key;
    unsigned long long best = 0ull; int best_slot = -1;
    for (int i = threadIdx.x; i < nl; i += blockDim.x) {
      const int s = live[i];
      const unsigned long long key = ((unsigned long long)(unsigned)cnts[s] << 32) | (unsigned long long)(0xFFFFFFFFu - firsts[s]);
      if (key < limit && (best_slot < 0 || key > best)) { best = key; best_slot = s; }
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
      // table exhausted: pad the rest
      for (int r = m + 1 + threadIdx.x; r < M; r += blockDim.x) { out_names[(long long)c * M + r] = -1; out_counts[(long long)c * M + r] = 0; }
      break;
    }
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
