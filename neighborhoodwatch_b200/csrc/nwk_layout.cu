// nwk_layout.cu -- layout conversion on the way into the engine.
//
// The reference's parquet files hold one column per dimension (generate_dataset.py:233-235), i.e.
// the vectors arrive column-major [d, rows]; the reference turns them into the row-major array knn()
// wants with cupy.copy(order='C') (cu_knn.py:211,235).  rows_from_columns is that step as one tiled
// shared-memory transpose.  HBM-bound: reads 4*rows*d bytes, writes 4*rows*d bytes.
//
// Tile = 64 dimensions x 64 rows, 256 threads.  A source column contributes 256 contiguous bytes per
// tile (16 lanes x float4), a destination row receives 256 contiguous bytes (16 lanes x float4): two
// full 128-byte lines per request on both sides, four independent 16-byte loads in flight per thread.
// (The first version moved 32 x 32 tiles with 4-byte accesses: one 128-byte line per column and tile,
// columns 4 MB apart -- 2.86 TB/s of read + write at 10^6 x 768, measured by scripts/bench_loader.py.)
// The float4 forms need 16-byte aligned bases and ld % 4 == 0 (reads), d % 4 == 0 (writes); other
// shapes take the scalar form of the same tile on that side.
#include "nwk_common.cuh"

namespace nwk {

constexpr int TR_TILE = 64;
constexpr int TR_THREADS = 256;
constexpr int TR_PITCH = TR_TILE + 1;      // odd pitch: column-wise and row-wise walks both spread over the banks

// src[c * ld + r] (c < d, r < rows)  ->  dst[r * d + c]
template <bool VIN, bool VOUT>
__global__ void __launch_bounds__(TR_THREADS)
rows_from_columns_kernel(const float* __restrict__ src, long long ld, int d, long long rows, float* __restrict__ dst) {
  __shared__ float tile[TR_TILE * TR_PITCH];          // tile[c][r]
  const long long tiles_r = (rows + TR_TILE - 1) / TR_TILE;
  const int tiles_c = (d + TR_TILE - 1) / TR_TILE;
  const long long ntiles = tiles_r * tiles_c;
  const int tx = threadIdx.x & 15;                     // 16 lanes x 4 elements along the contiguous side
  const int ty = threadIdx.x >> 4;                     // 16 lines per pass, 4 passes
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // consecutive CTAs take consecutive column tiles of the same row block: together they fill
    // whole destination rows
    const long long tr = t / tiles_c;
    const int tc = (int)(t - tr * tiles_c);
    const long long r0 = tr * TR_TILE;
    const int c0 = tc * TR_TILE;
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 16 * i;
      const long long r = r0 + 4 * tx;
      const float* p = src + (long long)c * ld + r;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        if (VIN && r + 3 < rows) {
          v[i] = __ldcs(reinterpret_cast<const float4*>(p));       // streamed: read once
        } else {
          if (r + 0 < rows) v[i].x = __ldcs(p + 0);
          if (r + 1 < rows) v[i].y = __ldcs(p + 1);
          if (r + 2 < rows) v[i].z = __ldcs(p + 2);
          if (r + 3 < rows) v[i].w = __ldcs(p + 3);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* s = tile + (ty + 16 * i) * TR_PITCH + 4 * tx;
      s[0] = v[i].x; s[1] = v[i].y; s[2] = v[i].z; s[3] = v[i].w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = r0 + ty + 16 * i;
      const int c = c0 + 4 * tx;
      const float* s = tile + (4 * tx) * TR_PITCH + (ty + 16 * i);
      const float4 w = make_float4(s[0], s[TR_PITCH], s[2 * TR_PITCH], s[3 * TR_PITCH]);
      if (r < rows) {
        float* q = dst + r * (long long)d + c;
        if (VOUT && c + 3 < d) {
          *reinterpret_cast<float4*>(q) = w;
        } else {
          if (c + 0 < d) q[0] = w.x;
          if (c + 1 < d) q[1] = w.y;
          if (c + 2 < d) q[2] = w.z;
          if (c + 3 < d) q[3] = w.w;
        }
      }
    }
    __syncthreads();
  }
}

int rows_from_columns(const float* src, long long ld, int d, long long rows, float* dst, cudaStream_t s) {
  if (rows == 0) return NWK_OK;
  const long long ntiles = ((rows + TR_TILE - 1) / TR_TILE) * ((d + TR_TILE - 1) / TR_TILE);
  long long blocks = ntiles;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const bool vin = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (ld & 3) == 0;
  const bool vout = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (d & 3) == 0;
  const unsigned g = (unsigned)blocks;
  if (vin && vout) rows_from_columns_kernel<true, true><<<g, TR_THREADS, 0, s>>>(src, ld, d, rows, dst);
  else if (vin) rows_from_columns_kernel<true, false><<<g, TR_THREADS, 0, s>>>(src, ld, d, rows, dst);
  else if (vout) rows_from_columns_kernel<false, true><<<g, TR_THREADS, 0, s>>>(src, ld, d, rows, dst);
  else rows_from_columns_kernel<false, false><<<g, TR_THREADS, 0, s>>>(src, ld, d, rows, dst);
  NWK_LAUNCH_CHECK();
  return NWK_OK;
}

}  // namespace nwk
