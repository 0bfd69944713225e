// Proxy for a single-pass E+M k-means kernel: 148 persistent CTAs stream 128-row tiles of X (12 "E-step" warps read tile t
// from HBM) while 4 "scatter" warps re-read tile t-1 (hopefully from L2) and red.global.add.v4.f32 its rows into
// sums[label].  Question: does the whole thing run near one HBM pass (60 us), or does the re-read miss L2?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fused_proxy tools/fused_em_proxy.cu && /tmp/fused_proxy
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {     // L2 evict_last-ish default vs evict_first
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int SCATTER_WARPS, bool DO_SCATTER>
__global__ void __launch_bounds__(512) proxy(const float* __restrict__ X, const int* __restrict__ labels, int N, int D,
                                             float* __restrict__ sums, float* __restrict__ sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (N + 127) / 128;
  constexpr int ES = 16 - SCATTER_WARPS;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int prev = -1;
  for (int t = blockIdx.x; ; t += gridDim.x) {
    const bool have = t < n_tiles;
    if (warp < ES) {
      if (have) {                                     // "E-step": stream the tile once, 128 rows x 768 floats
        for (int r = warp; r < 128; r += ES) {
          const long long row = (long long)t * 128 + r;
          if (row >= N) break;
          const float4* src = reinterpret_cast<const float4*>(X + row * D);
#pragma unroll
          for (int j = 0; j < 6; ++j) { const float4 v = ld_stream(src + j * 32 + lane); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        }
      }
    } else if (DO_SCATTER && prev >= 0) {             // "M-step" of the previous tile: re-read + scatter
      for (int r = warp - ES; r < 128; r += SCATTER_WARPS) {
        const long long row = (long long)prev * 128 + r;
        if (row >= N) break;
        const float4* src = reinterpret_cast<const float4*>(X + row * D);
        float* dst = sums + (size_t)labels[row] * D;
        float4 v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = ld_stream(src + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < 6; ++j) red_add_v4(dst + j * 128 + lane * 4, v[j]);
      }
    }
    __syncthreads();
    if (!have) break;
    prev = t;
  }
  if (acc.x == 123.456f) sink[0] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
  const int N = 127000, D = 768, K = 100;
  float* X; cudaMalloc(&X, (size_t)N * D * 4); cudaMemset(X, 0, (size_t)N * D * 4);
  int* lab; cudaMalloc(&lab, N * 4);
  std::vector<int> h(N);
  for (int i = 0; i < N; ++i) h[i] = rand() % K;
  cudaMemcpy(lab, h.data(), N * 4, cudaMemcpyHostToDevice);
  float* sums; cudaMalloc(&sums, (size_t)K * D * 4); cudaMemset(sums, 0, (size_t)K * D * 4);
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time_it = [&](const char* name, auto launch) {
    for (int it = 0; it < 3; ++it) launch();
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %7.1f us per pass (%5.0f GB/s of one X pass)  %s\n", name, ms * 100, (double)N * D * 4 / (ms / 10 * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  };
  time_it("stream only (16 warps)", [&] { proxy<0, false><<<148, 512>>>(X, lab, N, D, sums, sink); });
  time_it("stream only (12 warps)", [&] { proxy<4, false><<<148, 512>>>(X, lab, N, D, sums, sink); });
  time_it("stream + scatter of previous tile (4 warps)", [&] { proxy<4, true><<<148, 512>>>(X, lab, N, D, sums, sink); });
  time_it("stream + scatter of previous tile (8 warps)", [&] { proxy<8, true><<<148, 512>>>(X, lab, N, D, sums, sink); });
  time_it("stream(74 CTAs) + scatter (4 warps)", [&] { proxy<4, true><<<74, 512>>>(X, lab, N, D, sums, sink); });
  return 0;
}
