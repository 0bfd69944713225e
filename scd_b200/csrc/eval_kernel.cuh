// Evaluation either side of the naming round: the contingency matrix behind
//   gcd/project_utils/cluster_and_log_utils.py:47-49   for i in range(y_pred.size): w[y_pred[i], y_true[i]] += 1
//   main_unsup.py:149-167 evaluate_semantic_acc         (per-class match lists; the counts are w's cells)
// as one pass over the two label vectors.  w is [D, D] int64, row = predicted cluster, column = true class.
// `first_row[t]` = first row index whose true class is t (N when the class never occurs): the order in which
// Python's defaultdict meets the class names (main_unsup.py:152), which fixes the float summation order of
// semantic_acc_avg.  `col_masked[t]` = number of rows of true class t with mask set: split_cluster_acc_v2's
// old_classes_gt = set(y_true[mask]) / new_classes_gt = set(y_true[~mask]) (:43-44) are the classes with
// col_masked > 0 / column sum - col_masked > 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scd {

constexpr int kContSmemCells = 12288;          // 48 KB of int32 cells: D <= 110 takes the privatised path

// Labels may arrive as int64 or as float64 (the drivers build `targets` with np.append -> float64,
// main_unsup.py:118,132; split_cluster_acc_v2 casts with .astype(int) = truncation toward zero).
template <typename T>
__device__ __forceinline__ long long as_label(T v) { return (long long)v; }

template <typename TP, typename TT>
__global__ void __launch_bounds__(256)
contingency_kernel(const TP* __restrict__ y_pred, const TT* __restrict__ y_true, long long N, int D,
                   const unsigned char* __restrict__ mask, unsigned long long* __restrict__ w,
                   unsigned long long* __restrict__ first_row, unsigned long long* __restrict__ col_masked,
                   int* __restrict__ bad, int use_smem) {
  extern __shared__ int cont_sh[];
  const int cells = D * D;
  if (use_smem) {
    for (int c = threadIdx.x; c < cells; c += blockDim.x) cont_sh[c] = 0;
    __syncthreads();
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const long long p = as_label(y_pred[i]);
    const long long t = as_label(y_true[i]);
    if (p < 0 || p >= D || t < 0 || t >= D) { atomicOr(bad, 1); continue; }
    if (use_smem) atomicAdd(&cont_sh[(int)p * D + (int)t], 1);
    else atomicAdd(&w[p * D + t], 1ull);
    if (first_row) atomicMin(&first_row[t], (unsigned long long)i);
    if (mask && mask[i]) atomicAdd(&col_masked[t], 1ull);
  }
  if (use_smem) {
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x)
      if (cont_sh[c]) atomicAdd(&w[c], (unsigned long long)cont_sh[c]);
  }
}

__global__ void fill_u64_kernel(unsigned long long* __restrict__ p, long long n, unsigned long long v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace scd
