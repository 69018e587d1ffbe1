// nwk_synth.cu -- deterministic synthetic vectors, bit-identical to neighborhoodwatch_b200/synth.py.
//
// Fixture style of the reference's tests (tests/test_02_cu_knn.py:21-28 and
// nw_utils.py:56-59): every component uniform in [0,1), rows L2-normalised.  Here the
// uniform draw is a counter-based hash of (seed, row, col) so any slab of a 10^8-row base
// can be regenerated on either side, and the normalisation is done in integer / fp64
// arithmetic so numpy and CUDA agree bit for bit:
//     u(row,col) = splitmix64(seed*GOLDEN + row*d + col) >> 40            (24 bits)
//     x = u * 2^-24;  S = sum u^2 (exact in uint64);  y = float(double(u) / sqrt(double(S)))
#include "nwk_common.cuh"

namespace nwk {

__device__ __forceinline__ unsigned int synth_u24(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed * 0x9E3779B97F4A7C15ull + idx;
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 40);
}

__global__ void __launch_bounds__(256)
synth_kernel(float* __restrict__ out, long long rows, int d, unsigned long long seed, long long row0,
             int normalize) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    const unsigned long long base_idx = (unsigned long long)(row0 + r) * (unsigned long long)d;
    double inv = 0.0;
    if (normalize) {
      unsigned long long s = 0;
      for (int c = lane; c < d; c += 32) {
        unsigned long long u = synth_u24(seed, base_idx + c);
        s += u * u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      inv = sqrt((double)s);
    }
    for (int c = lane; c < d; c += 32) {
      unsigned int u = synth_u24(seed, base_idx + c);
      float v = normalize ? (float)((double)u / inv) : (float)u * 5.9604644775390625e-08f;
      out[r * d + c] = v;
    }
  }
}

int synth_fill(float* out, long long rows, int d, unsigned long long seed, long long row0, int normalize,
               cudaStream_t s) {
  if (rows == 0) return NWK_OK;
  long long blocks = (rows + 7) / 8;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  synth_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, rows, d, seed, row0, normalize);
  NWK_LAUNCH_CHECK();
  return NWK_OK;
}

}  // namespace nwk
