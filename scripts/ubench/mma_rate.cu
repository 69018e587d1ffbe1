// Microbenchmark: tcgen05.mma (kind::f16, M=128, cta_group::1) throughput for short accumulation chains.
// One CTA per SM; one thread issues `tiles` accumulator tiles of nkb K-blocks (4 MMAs of K=16 each) from
// operands that sit in shared memory (garbage bits), alternating between accumulator stages, one commit
// per K-block and per tile like the filter kernel.  Optionally a second thread streams `tma_kb` KB per
// K-block from global memory into another part of shared memory with cp.async.bulk (shared-memory write
// traffic of a TMA producer, without any dependency on the MMAs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I neighborhoodwatch_b200/csrc -o scripts/ubench/mma_rate.bin scripts/ubench/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "nwk_ptx.cuh"
using namespace nwk;

__global__ void __launch_bounds__(128, 1) k(int tiles, int nkb, int N, int nacc, int tma_kb, const unsigned char* src,
                                            long long* clk) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_done, bar_tile[4], bar_tma[2];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_done, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar_tile[i], 1);
    ptx::mbar_init(&bar_tma[0], 1); ptx::mbar_init(&bar_tma[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(&s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, N);
    const long long t0 = clock64();
    int acc = 0, stage = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d_tmem = tmem_base + acc * N;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t sa = ptx::smem_u32(smem + stage * 48 * 1024);
        const uint64_t da = ptx::umma_desc_sw128(sa);
        const uint64_t db = ptx::umma_desc_sw128(sa + 16 * 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ptx::mma_f16_ss(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
        ptx::tc_commit(&bar_tile[2 + (stage & 1)]);        // stands in for the operand-stage release
        if (++stage == 2) stage = 0;
      }
      ptx::tc_commit(&bar_tile[acc & 1]);
      if (++acc == nacc) acc = 0;
    }
    ptx::tc_commit(&bar_done);
    while (!ptx::mbar_try_wait(&bar_done, 0)) {}
    clk[blockIdx.x] = clock64() - t0;
  } else if (warp == 0 && lane == 0 && tma_kb > 0) {
    // free-running bulk copies into the upper 96 KB: 2 in flight
    uint32_t ph[2] = {0, 0};
    const int total = tiles * nkb;
    unsigned char* dst = smem + 96 * 1024;
    for (int i = 0; i < total + 2; ++i) {
      const int b = i & 1;
      if (i >= 2) { while (!ptx::mbar_try_wait(&bar_tma[b], ph[b])) {} ph[b] ^= 1; }
      if (i < total) {
        ptx::mbar_arrive_expect_tx(&bar_tma[b], tma_kb * 1024);
        ptx::bulk_load_1d(dst + b * 48 * 1024, src + ((size_t)(blockIdx.x * 7 + i) % 2048) * 49152, tma_kb * 1024, &bar_tma[b]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// The same MMAs fed through a 4-stage full / empty mbarrier ring by a producer warp (bulk copies of
// `stage_kb` KB per K-block from an L2-resident buffer), both roles as converged warps with elect.sync:
// the filter kernel's operand pipeline without its epilogue and schedule.
__global__ void __launch_bounds__(128, 1) kring(int tiles, int nkb, int stage_kb, const unsigned char* src, long long* clk) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[4], empty[4], bar_done, bar_tile[2];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&bar_done, 1); ptx::mbar_init(&bar_tile[0], 1); ptx::mbar_init(&bar_tile[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(&s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const int total = tiles * nkb;
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < total; ++i) {
      while (!ptx::mbar_try_wait(&empty[stage], phase ^ 1)) {}
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&full[stage], stage_kb * 1024);
        ptx::bulk_load_1d(smem + stage * 48 * 1024, src + ((size_t)(blockIdx.x * 7 + i) % 2048) * 49152, stage_kb * 1024, &full[stage]);
      }
      __syncwarp();
      if (++stage == 4) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::umma_idesc_f16(128, 256);
    const long long t0 = clock64();
    int stage = 0; uint32_t phase = 0; int acc = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < nkb; ++kb) {
        while (!ptx::mbar_try_wait(&full[stage], phase)) {}
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * 48 * 1024);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db = ptx::umma_desc_sw128(sa + 16 * 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::mma_f16_ss(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
          ptx::tc_commit(&empty[stage]);
          if (kb == nkb - 1) ptx::tc_commit(&bar_tile[acc]);
        }
        __syncwarp();
        if (++stage == 4) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
    }
    if (ptx::elect_one()) ptx::tc_commit(&bar_done);
    __syncwarp();
    while (!ptx::mbar_try_wait(&bar_done, 0)) {}
    if (lane == 0) clk[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* clk; cudaMalloc(&clk, sms * 8);
  unsigned char* src; cudaMalloc(&src, (size_t)2048 * 49152 + 65536); cudaMemset(src, 0x3c, (size_t)2048 * 49152 + 65536);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int nkb, N, nacc, tma_kb; } cfgs[] = {
      {2, 256, 2, 0}, {4, 256, 2, 0}, {12, 256, 2, 0}, {24, 256, 2, 0}, {2, 128, 4, 0}, {2, 256, 2, 32}, {2, 256, 2, 48}, {12, 256, 2, 48}, {2, 128, 4, 32}};
  for (auto c : cfgs) {
    const int tiles = 24000 / c.nkb;
    for (int grid : {1, sms}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, 128, smem>>>(tiles, c.nkb, c.N, c.nacc, c.tma_kb, src, clk);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      const double mmas = (double)tiles * c.nkb * 4;
      printf("nkb %2d N %3d accs %d tma %2d KB/kblock grid %3d: %.1f clk per MMA (ideal %d), %.0f clk per tile, %.2f GHz, %.0f TFLOP/s chip\n",
             c.nkb, c.N, c.nacc, c.tma_kb, grid, h / mmas, c.N / 2, h / (double)tiles, h / (ms * 1e6),
             grid * mmas * 128.0 * c.N * 16 * 2 / (ms * 1e-3) / 1e12);
    }
  }
  cudaFuncSetAttribute(kring, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int nkb : {2, 12})
    for (int kb : {16, 32, 48})
      for (int grid : {1, sms}) {
        const int tiles = 24000 / nkb;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          kring<<<grid, 128, smem>>>(tiles, nkb, kb, src, clk);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaEventElapsedTime(&ms, e0, e1);
        }
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("ring: nkb %2d, %2d KB per K-block, grid %3d: %.0f clk per tile (MMAs alone %d), %.2f GHz, %.0f GB/s per SM\n", nkb, kb,
               grid, h / (double)tiles, nkb * 512, h / (ms * 1e6), (double)tiles * nkb * kb * 1024 / (ms * 1e-3) / 1e9);
      }
  return 0;
}
