// nwk_common.cuh -- shared host/device helpers for libnwk (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/nwk.h"

namespace nwk {

// ---------------------------------------------------------------------------------------
// error plumbing: thread-local message + status codes, no exceptions across the ABI
// ---------------------------------------------------------------------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define NWK_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::nwk::fail(NWK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

#define NWK_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    ::nwk::count_launch();                                                               \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ::nwk::fail(NWK_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

#define NWK_TRY(expr)                                                                    \
  do {                                                                                   \
    int _s = (expr);                                                                     \
    if (_s != NWK_OK) return _s;                                                         \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  size_t size;
  size_t off = 0;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base ? base + off : nullptr);
    off += count * sizeof(T);
    return p;
  }
  bool ok() const { return off <= size; }
};

// Per-device state.  Function attributes (dynamic shared-memory opt-ins), the SM count and
// the library's own streams are properties of ONE device, so everything the library caches
// is keyed by the ordinal of the device that is current when an entry point is called: a
// process may drive several GPUs (one host thread per device, nwk_mg_*) or one (the
// one-process-per-GPU model of the Python drivers).
constexpr int NWK_MAX_DEVICES = 64;
constexpr int NWK_SIDE_STREAMS = 2;
struct DevState {
  int checked = 0;              // 1 once the device was verified to be sm_100
  int sms = 0;
  int tc_optin = 0, seg_optin = 0, scan_optin = 0, ms_optin = 0;
  int rr_optin = 48 * 1024, merge_rank_optin = 48 * 1024, ms_ex_optin = 48 * 1024;
  // fork / join inside one call (nwk_tc.cu): side streams and the events that order them
  cudaStream_t side[NWK_SIDE_STREAMS] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[NWK_SIDE_STREAMS] = {nullptr, nullptr};
  cudaEvent_t ev_chunk[16] = {};
  std::recursive_mutex mu;      // one call at a time enqueues through the side streams / events of a device
};
int device_check();             // verifies that the current device is sm_100; caches the answer per device
DevState* dev_state();          // state of the current device (device_check() has succeeded on it)
int sm_count();

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
// Order-preserving map float -> uint32 (smaller float <=> smaller uint).  NaN maps above +inf.
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// (key, id) pair ordered lexicographically; +inf keys with id = INT64_MAX are padding
struct KI {
  double key;
  long long id;
};
__device__ __forceinline__ bool ki_less(double ka, long long ia, double kb, long long ib) {
  return (ka < kb) || (ka == kb && ia < ib);
}

// In-place bitonic sort of n (power of two) (key,id) pairs held in shared memory by a whole
// CTA.  Ascending by (key, id).  All threads of the block must call it.
__device__ __forceinline__ void block_bitonic_sort(double* keys, long long* ids, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        int hi = lo | stride;
        bool up = ((lo & size) == 0);
        double ka = keys[lo], kb = keys[hi];
        long long ia = ids[lo], ib = ids[hi];
        bool swap = up ? ki_less(kb, ib, ka, ia) : ki_less(ka, ia, kb, ib);
        if (swap) {
          keys[lo] = kb; keys[hi] = ka;
          ids[lo] = ib; ids[hi] = ia;
        }
      }
    }
  }
  __syncthreads();
}

// same for packed 64-bit keys (ordered-float << 32 | row id): used for the candidate logs
__device__ __forceinline__ void block_bitonic_sort_u64(unsigned long long* v, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        int hi = lo | stride;
        bool up = ((lo & size) == 0);
        unsigned long long a = v[lo], b = v[hi];
        if (up ? (b < a) : (a < b)) { v[lo] = b; v[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// float -> double is exact, so it can be done on the integer pipe: F2F.F64.F32 shares the fp64 pipe with
// the DFMAs and costs about four of them (measured on the MaxSim re-scoring kernel: a variant with twice the
// conversions per DFMA ran 1.33 x slower), while the integer instructions overlap with them.  Zeros,
// denormals, inf and nan take the conversion instruction.  Used by every exact fp64 path that reads fp32
// inputs; being exact, it changes no result bit.
__device__ __forceinline__ double f2d_exact(float f) {
  const uint32_t b = __float_as_uint(f);
  const uint32_t a = b & 0x7fffffffu;
  if (a - 0x00800000u < 0x7f000000u)                     // a normal number
    return __hiloint2double((int)((b & 0x80000000u) | ((a >> 3) + 0x38000000u)), (int)(b << 29));
  return (double)f;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// metric finalisation in fp64 -- the one definition every exact path uses
//   sq-L2 is accumulated directly as sum (q-b)^2 by the callers; this covers dot / cosine
__device__ __forceinline__ double finish_metric(int metric, double dot, double qn2, double bn2) {
  if (metric == NWK_METRIC_DOT) return 1.0 - dot;
  return 1.0 - dot / (sqrt(qn2) * sqrt(bn2));
}

}  // namespace nwk
