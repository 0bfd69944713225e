// Micro-benchmark: can the M-step be a per-row scatter of red.global.add.v4.f32 into sums[label] (no sort)?  That is what a
// single-pass E+M kernel would do with its tile's rows right after the argmin.  Measures 127000 x 768 fp32 rows, random labels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_bench tools/red_scatter_bench.cu && /tmp/red_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp per `rows_per_warp` consecutive rows; `agg` consecutive rows are summed in registers before the reds when they
// share a label (labels are generated so that runs of `agg` rows share one)
template <int NV>
__global__ void scatter_kernel(const float* __restrict__ X, const int* __restrict__ labels, int N, int D, float* __restrict__ sums,
                               int rows_per_warp, int agg) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int r0 = warp * rows_per_warp;
  for (int r = r0; r < min(r0 + rows_per_warp, N); r += agg) {
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < agg && r + a < N; ++a) {
      const float* src = X + (size_t)(r + a) * D;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + j * 128 + lane * 4));
        acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
      }
    }
    float* dst = sums + (size_t)labels[r] * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) red_add_v4(dst + j * 128 + lane * 4, acc[j]);
  }
}

int main() {
  const int N = 127000, D = 768;
  float* X; cudaMalloc(&X, (size_t)N * D * 4); cudaMemset(X, 0, (size_t)N * D * 4);
  int* lab; cudaMalloc(&lab, N * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int K : {100, 1000}) {
    float* sums; cudaMalloc(&sums, (size_t)K * D * 4); cudaMemset(sums, 0, (size_t)K * D * 4);
    for (int agg : {1, 2, 4, 8, 32}) {
      std::vector<int> h(N);
      for (int i = 0; i < N; ++i) h[i] = (i % agg == 0) ? rand() % K : h[i - 1];
      cudaMemcpy(lab, h.data(), N * 4, cudaMemcpyHostToDevice);
      for (int rpw : {8, 32}) {
        if (rpw < agg) continue;
        const int warps = (N + rpw - 1) / rpw, blocks = (warps * 32 + 255) / 256;
        for (int it = 0; it < 3; ++it) scatter_kernel<6><<<blocks, 256>>>(X, lab, N, D, sums, rpw, agg);
        cudaEventRecord(e0);
        for (int it = 0; it < 10; ++it) scatter_kernel<6><<<blocks, 256>>>(X, lab, N, D, sums, rpw, agg);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("K=%4d rows/label-run=%2d rows/warp=%2d : %7.1f us per pass (%5.0f GB/s of X, %5.1f M red.v4)  %s\n", K, agg, rpw, ms * 100,
               (double)N * D * 4 / (ms / 10 * 1e-3) / 1e9, (double)N / agg * D / 4 / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
    cudaFree(sums);
  }
  return 0;
}
