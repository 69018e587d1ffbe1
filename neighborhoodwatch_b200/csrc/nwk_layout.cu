// nwk_layout.cu -- layout conversion on the way into the engine.
//
// The reference's parquet files hold one column per dimension (generate_dataset.py:233-235), i.e.
// the vectors arrive column-major [d, rows]; the reference turns them into the row-major array knn()
// wants with cupy.copy(order='C') (cu_knn.py:211,235).  rows_from_columns is that step as one tiled
// shared-memory transpose: 32 x 32 element tiles, both the reads (along rows of the source) and the
// writes (along rows of the destination) are full 128-byte lines.  HBM-bound: reads 4*rows*d bytes,
// writes 4*rows*d bytes.
#include "nwk_common.cuh"

namespace nwk {

constexpr int TR_TILE = 32;
constexpr int TR_ROWS = 8;        // thread rows per tile: each thread moves TR_TILE / TR_ROWS elements

// src[c * ld + r] (c < d, r < rows)  ->  dst[r * d + c]
__global__ void __launch_bounds__(TR_TILE * TR_ROWS)
rows_from_columns_kernel(const float* __restrict__ src, long long ld, int d, long long rows, float* __restrict__ dst) {
  __shared__ float tile[TR_TILE][TR_TILE + 1];
  const long long tiles_r = (rows + TR_TILE - 1) / TR_TILE;
  const int tiles_c = (d + TR_TILE - 1) / TR_TILE;
  const long long ntiles = tiles_r * tiles_c;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // consecutive CTAs take consecutive column tiles of the same row block: their writes fill
    // whole destination rows
    const long long tr = t / tiles_c;
    const int tc = (int)(t - tr * tiles_c);
    const long long r0 = tr * TR_TILE;
    const int c0 = tc * TR_TILE;
#pragma unroll
    for (int j = threadIdx.y; j < TR_TILE; j += TR_ROWS) {
      const int c = c0 + j;
      const long long r = r0 + threadIdx.x;
      tile[j][threadIdx.x] = (c < d && r < rows) ? __ldg(src + (long long)c * ld + r) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = threadIdx.y; j < TR_TILE; j += TR_ROWS) {
      const long long r = r0 + j;
      const int c = c0 + threadIdx.x;
      if (r < rows && c < d) dst[r * d + c] = tile[threadIdx.x][j];
    }
    __syncthreads();
  }
}

int rows_from_columns(const float* src, long long ld, int d, long long rows, float* dst, cudaStream_t s) {
  if (rows == 0) return NWK_OK;
  const long long ntiles = ((rows + TR_TILE - 1) / TR_TILE) * ((d + TR_TILE - 1) / TR_TILE);
  long long blocks = ntiles;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  rows_from_columns_kernel<<<(unsigned)blocks, dim3(TR_TILE, TR_ROWS), 0, s>>>(src, ld, d, rows, dst);
  NWK_LAUNCH_CHECK();
  return NWK_OK;
}

}  // namespace nwk
