// Why is ONE warp per SM slow at "re-read a row, reduce it into sums[label]"?  148 CTAs x 1 warp, S-deep cp.async ring.
//   variants: full | copies only (no red) | reds only (row kept in registers) | copies + plain stores instead of red
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mwarp tools/mwarp_bench.cu && /tmp/mwarp
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <stdint.h>

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void red4(float* a, float4 v) { asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

template <int MODE, int S, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k(const float* __restrict__ X, const int* __restrict__ lab, int N, float* __restrict__ sums, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* my = sm + warp * S * 3072;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(my);
  float4 keep = make_float4(0, 0, 0, 0);
  const int rows_per = (N + gridDim.x * WARPS - 1) / (gridDim.x * WARPS);
  const int r0 = (blockIdx.x * WARPS + warp) * rows_per, r1 = min(r0 + rows_per, N);
  int issued = r0;
  auto issue = [&](int r) {
    if (MODE == 2) return;
    const uint32_t dst = base + ((r - r0) % S) * 3072;
    const float* src = X + (size_t)r * 768;
#pragma unroll
    for (int j = 0; j < 6; ++j) cp_async_16(dst + (j * 128 + lane * 4) * 4, src + j * 128 + lane * 4);
    cp_commit();
  };
  for (; issued < min(r0 + S, r1); ++issued) issue(issued);
  for (int r = r0; r < r1; ++r) {
    if (MODE != 2) { if (r1 - r >= S) cp_wait<S - 1>(); else cp_wait<0>(); }
    const float* src = reinterpret_cast<const float*>(my + ((r - r0) % S) * 3072);
    float* dst = sums + (size_t)lab[r] * 768;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float4 v = MODE == 2 ? make_float4(1.f, 2.f, 3.f, (float)r) : *reinterpret_cast<const float4*>(src + j * 128 + lane * 4);
      if (MODE == 0 || MODE == 2) red4(dst + j * 128 + lane * 4, v);
      else if (MODE == 3) *reinterpret_cast<float4*>(dst + j * 128 + lane * 4) = v;
      else { keep.x += v.x; keep.y += v.y; keep.z += v.z; keep.w += v.w; }
    }
    if (issued < r1) { issue(issued); ++issued; }
  }
  if (keep.x == 123.f) sink[0] = keep.x + keep.y + keep.z + keep.w;
}

int main() {
  const int N = 127000, D = 768, K = 100;
  float* X; cudaMalloc(&X, (size_t)N * D * 4); cudaMemset(X, 0, (size_t)N * D * 4);
  int* lab; cudaMalloc(&lab, N * 4);
  std::vector<int> h(N); for (int i = 0; i < N; ++i) h[i] = rand() % K;
  cudaMemcpy(lab, h.data(), N * 4, cudaMemcpyHostToDevice);
  float* sums; cudaMalloc(&sums, (size_t)K * D * 4); cudaMemset(sums, 0, (size_t)K * D * 4);
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, int warps, int S) {
    const int smem = warps * S * 3072;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int i = 0; i < 2; ++i) kern<<<148, 32 * warps, smem>>>(X, lab, N, sums, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) kern<<<148, 32 * warps, smem>>>(X, lab, N, sums, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-52s %8.1f us per pass  (%6.1f ns per row per SM)  %s\n", name, ms * 200, ms / 5 * 1e6 / (N / 148.0), cudaGetErrorString(cudaGetLastError()));
  };
  run("1 warp/SM, S=10: copies + red.v4", k<0, 10, 1>, 1, 10);
  run("1 warp/SM, S=10: copies only", k<1, 10, 1>, 1, 10);
  run("1 warp/SM      : red.v4 only (registers)", k<2, 10, 1>, 1, 10);
  run("1 warp/SM, S=10: copies + plain st.v4", k<3, 10, 1>, 1, 10);
  run("1 warp/SM, S=16: copies + red.v4", k<0, 16, 1>, 1, 16);
  run("2 warps/SM, S=10: copies + red.v4", k<0, 10, 2>, 2, 10);
  run("4 warps/SM, S=10: copies + red.v4", k<0, 10, 4>, 4, 10);
  run("4 warps/SM      : red.v4 only", k<2, 10, 4>, 4, 10);
  run("8 warps/SM, S=6: copies + red.v4", k<0, 6, 8>, 8, 6);
  return 0;
}
