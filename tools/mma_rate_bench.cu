// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 in, fp32 accumulate) as a function of the tile width N, of the
// A-operand source (shared memory descriptor vs tensor memory) and of cta_group (1: M = 128, 2: M = 256 over a CTA pair).
// One thread issues `iters` dependent-free MMAs into the same accumulator and commits once; operands are whatever the
// (zeroed) shared memory holds - only the timing matters.  Used to pick the naming kernel's tile shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_rate_bench tools/mma_rate_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#include "../scd_b200/csrc/ptx.cuh"

using namespace scd;

struct Params {
  int n, iters, ts, swizzle_bytes, n_acc, a_cols, feeders, tmem_readers, random_data;
  const char* src;
  long long* cycles;
};

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t bar = sbase + 96 * 1024, tptr = bar + 16;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) {
    // zeros, or pseudo-random bf16 pairs in [-1, 1) (exponent 0x3f / 0xbf region) - operand toggling changes the power draw
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t lo = 0x3f00u | (h & 0x80ffu), hi = 0x3f00u | ((h >> 16) & 0x80ffu);
    reinterpret_cast<uint32_t*>(smem)[i] = p.random_data ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init_cluster(); *reinterpret_cast<volatile int*>(smem + 96 * 1024 + 64) = 0; }
  if (warp == 1) { ptx::tmem_alloc<CG>(tptr, 512); ptx::tmem_relinquish<CG>(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before_sync();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + 96 * 1024 + 16);
  const bool issuer = threadIdx.x == 0 && (CG == 1 || ptx::cluster_ctarank() == 0);
  if (issuer) {
    const uint32_t idesc = ptx::make_idesc_bf16_f32(CG == 2 ? 256 : 128, (uint32_t)p.n);
    const uint64_t adesc = ptx::make_kmajor_desc(sbase, 128);
    const uint64_t bdesc = ptx::make_kmajor_desc(sbase + 32 * 1024, (uint32_t)p.swizzle_bytes);
    // descriptors of the 4 k-steps of a 64-wide k-block and the accumulator of each slot are loop invariant: the loop
    // body is four back-to-back MMAs, nothing else
    uint64_t ad[4], bd[4]; uint32_t at[4], dd[4];
    for (int j = 0; j < 4; ++j) {
      ad[j] = adesc + (uint64_t)(j * 2); bd[j] = bdesc + (uint64_t)(j * 2);
      at[j] = tmem_base + p.a_cols + j * 8;
      dd[j] = tmem_base + (uint32_t)(j % p.n_acc) * (uint32_t)p.n;
    }
    const long long t0 = clock64();
    if (p.ts) {
      for (int i = 0; i < p.iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_bf16_ts<CG>(dd[j], at[j], bd[j], idesc, 1u);
      }
    } else {
      for (int i = 0; i < p.iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::umma_bf16<CG>(dd[j], ad[j], bd[j], idesc, 1u);
      }
    }
    ptx::umma_commit<CG>(bar, 0b11);
    ptx::mbar_wait(bar, 0, 1);
    p.cycles[blockIdx.x] = clock64() - t0;
  }
  // background traffic while the issuer runs: `feeders` warps stream 12 KB bulk copies global -> shared memory (what the
  // B ring's TMA does), `tmem_readers` warps read accumulator columns with tcgen05.ld (what the epilogue does)
  volatile int* done_flag = reinterpret_cast<volatile int*>(smem + 96 * 1024 + 64);
  if (warp >= 2 && (int)warp - 2 < p.feeders && (threadIdx.x & 31) == 0) {
    const uint32_t fb = sbase + 96 * 1024 + 128 + 16 * (warp - 2);
    const uint32_t dst = sbase + 64 * 1024 + (warp - 2) * 12288;
    ptx::mbar_init(fb, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t ph = 0; long long moved = 0;
    while (!*done_flag) {
      ptx::mbar_arrive_expect_tx(fb, 12288);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(p.src + ((moved * 12288) & ((32ll << 20) - 1))), "r"(12288), "r"(fb) : "memory");
      ptx::mbar_wait(fb, ph, 2); ph ^= 1; ++moved;
    }
    p.cycles[148 + blockIdx.x] = moved;
  }
  if (warp >= 2 && (int)warp - 2 >= p.feeders && (int)warp - 2 < p.feeders + p.tmem_readers) {
    uint32_t r[32]; uint32_t acc = 0;
    while (!*done_flag) {
      ptx::tmem_ld_32x32(tmem_base + ((warp & 3u) * 32u << 16) + 0, r);
      ptx::tmem_ld_wait(r);
      acc += r[0];
    }
    if (acc == 0x12345678u) p.cycles[0] = 0;
  }
  if (issuer) { *done_flag = 1; }
  if (threadIdx.x == 0 && !issuer) { ptx::mbar_wait(bar, 0, 3); *done_flag = 1; }     // peer CTA: the commit is multicast to its barrier too
  ptx::tc_fence_before_sync();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<CG>(tmem_base, 512);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 296 * 8);
  const int smem = 98 * 1024 + 1024;
  cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  char* src; cudaMalloc(&src, 33ll << 20); cudaMemset(src, 0, 33ll << 20);
  printf("%-10s %-4s %5s %7s %7s %12s %10s %14s\n", "cta_group", "A", "N", "feeders", "random", "cyc/MMA", "ideal", "feed B/clk/SM");
  for (int cg : {2}) {
    for (int ts : {0, 1}) {
      for (int n : {128, 192, 256}) {
        for (int mode = 0; mode < 4; ++mode) {
          const int feeders = mode >= 2 ? 2 : 0, readers = 0, random_data = mode & 1;
          Params p{n, 65536, ts, 128, 1, 480, feeders, readers, random_data, src, cyc};
          cudaError_t e;
          for (int rep = 0; rep < 2; ++rep) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, mma_kernel<2>, p);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cg=%d ts=%d n=%d failed: %s\n", cg, ts, n, cudaGetErrorString(e)); return 1; }
          }
          std::vector<long long> h(296);
          cudaMemcpy(h.data(), cyc, 296 * 8, cudaMemcpyDeviceToHost);
          double mean = 0; int cnt = 0;
          for (int i = 0; i < 148; i += cg) { mean += (double)h[i]; ++cnt; }
          mean /= cnt;
          const double fed = feeders ? (double)h[148] * 12288.0 * feeders / mean : 0.0;
          fflush(stdout);
          printf("%-10d %-4s %5d %7d %7d %12.1f %10.1f %14.1f\n", cg, ts ? "tmem" : "smem", n, feeders, random_data, mean / p.iters, 128.0 * n / 256.0, fed);
        }
      }
    }
  }
  return 0;
}
