// Micro-benchmark behind the naming kernel's operand-feed design: how many bytes per clock per SM can TMA
// pull out of L2 when every SM streams the same 32 MB vocabulary [V, 768] bf16, as a function of the box shape and of
// the bytes in flight (ring depth).  One thread per CTA runs the ring: wait for the oldest box, re-issue into its slot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_feed_bench tools/tma_feed_bench.cu && ./tma_feed_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../scd_b200/csrc/ptx.cuh"

using namespace scd;

struct Params {
  int box_rows, box_k, stages, iters, tiles_rows, kblocks, hold, producers, hint, prefetch, bulk1d;
  const void* base;
  long long* cycles;     // [grid]
};

__global__ void __launch_bounds__(128, 1) feed_kernel(const __grid_constant__ CUtensorMap map, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t bytes = (uint32_t)p.box_rows * p.box_k * 2;
  const uint32_t bar0 = sbase + 200 * 1024;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages * p.producers; ++s) ptx::mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.prefetch) ptx::prefetch_tensormap(&map);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && warp < p.producers) {
    const long long t0 = clock64();
    const uint32_t ring = sbase + warp * p.stages * bytes, bars = bar0 + 8 * warp * p.stages;
    const uint64_t hint = p.hint == 0 ? ptx::kEvictNormal : (p.hint == 1 ? ptx::kEvictLast : ptx::kEvictFirst);
    // every CTA starts at a different row tile (the kernel's staggered sweep) and walks k-blocks inside a tile first;
    // producer w of a CTA takes every producers-th box
    int tile = (int)(((long long)blockIdx.x * p.tiles_rows) / gridDim.x), kb = warp;
    while (kb >= p.kblocks) { kb -= p.kblocks; ++tile; }
    for (int i = 0; i < p.iters + p.stages; ++i) {
      const int s = i % p.stages;
      if (i >= p.stages) {
        ptx::mbar_wait(bars + 8 * s, ((i / p.stages) - 1) & 1, 900 + s);
        if (p.hold > 0) { const long long c = clock64(); while (clock64() - c < p.hold) {} }
      }
      if (i < p.iters) {
        ptx::mbar_arrive_expect_tx(bars + 8 * s, bytes);
        if (p.bulk1d) {
          // contiguous pre-tiled image: box index (tile, kb) -> bytes at ((tile * kblocks) + kb) * bytes
          const char* src = (const char*)p.base + ((size_t)tile * p.kblocks + kb) * bytes;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                       ::"r"(ring + s * bytes), "l"(src), "r"(bytes), "r"(bars + 8 * s), "l"(hint) : "memory");
        } else {
          ptx::tma_load_2d<1>(ring + s * bytes, &map, bars + 8 * s, kb * p.box_k, tile * p.box_rows, hint);
        }
        kb += p.producers;
        while (kb >= p.kblocks) { kb -= p.kblocks; if (++tile >= p.tiles_rows) tile = 0; }
      }
    }
    if (warp == 0) p.cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long V = 21000, D = 768;
  __nv_bfloat16* W;
  cudaMalloc(&W, V * D * 2);
  cudaMemset(W, 0, V * D * 2);
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
  struct Shape { int rows, k; CUtensorMapSwizzle sw; const char* name; };
  Shape shapes[] = {{128, 32, CU_TENSOR_MAP_SWIZZLE_64B, "128x32 SW64 (8 KB)"},
                    {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128x64 SW128 (16 KB)"},
                    {256, 64, CU_TENSOR_MAP_SWIZZLE_128B, "256x64 SW128 (32 KB)"},
                    {64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "64x64 SW128 (8 KB)"},
                    {128, 16, CU_TENSOR_MAP_SWIZZLE_32B, "128x16 SW32 (4 KB)"}};
  printf("%-22s %5s %6s %5s %5s %5s %10s %10s %12s %10s\n", "box", "prod", "stages", "hint", "pref", "hold", "inflightKB", "B/clk/SM", "chip B/clk", "cyc/box");
  for (const Shape& sh : shapes) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)V};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {(cuuint32_t)sh.k, (cuuint32_t)sh.rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sh.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int bytes = sh.rows * sh.k * 2;
    struct Var { int producers, stages, hint, prefetch, hold, bulk1d; };
    Var vars[] = {{1, 1, 1, 1, 0, 0}, {1, 2, 1, 1, 0, 0}, {1, 4, 1, 1, 0, 0}, {2, 4, 1, 1, 0, 0}, {4, 2, 1, 1, 0, 0}, {1, 4, 1, 1, 1, 0}, {2, 4, 1, 1, 1, 0},
                  {1, 1, 1, 1, 0, 1}, {1, 2, 1, 1, 0, 1}, {1, 4, 1, 1, 0, 1}, {1, 8, 1, 1, 0, 1}, {2, 4, 1, 1, 0, 1}, {4, 2, 1, 1, 0, 1}, {1, 4, 1, 1, 1, 1}, {2, 4, 1, 1, 1, 1}};
    for (const Var& v : vars) {
      if ((long long)v.stages * v.producers * bytes > 192 * 1024) continue;
      Params p;
      p.box_rows = sh.rows; p.box_k = sh.k; p.stages = v.stages; p.producers = v.producers; p.hint = v.hint; p.prefetch = v.prefetch; p.bulk1d = v.bulk1d; p.base = W;
      p.iters = (int)(48ll * 1024 * 1024 / bytes / v.producers);
      p.tiles_rows = (int)(V / sh.rows); p.kblocks = (int)(D / sh.k);
      p.hold = v.hold ? bytes / 32 : 0;                  // hold each slot for the time the MMA needs it (32 B/clk)
      p.cycles = cyc;
      for (int rep = 0; rep < 2; ++rep) {
        feed_kernel<<<148, 128, 222 * 1024, 0>>>(map, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (long long c : h) mean += (double)c;
      mean /= 148;
      const double bpc = (double)p.iters * v.producers * bytes / mean;
      printf("%-22s %s %5d %6d %5d %5d %5d %10.0f %10.2f %12.0f %10.0f\n", sh.name, v.bulk1d ? "bulk1d" : "tensor", v.producers, v.stages, v.hint, v.prefetch, p.hold,
             v.stages * v.producers * bytes / 1024.0, bpc, bpc * 148, mean / p.iters);
    }
  }
  return 0;
}
