// nwk_xchg.cu -- the cross-GPU reduction over peer memory: gather, k-way merge and broadcast of the
// result without a collective library call.
//
// Multi-GPU step (one process per GPU): every rank produces its exact local top-k [Q, k] with global
// ids, then the W lists of a query are merged (merge.py:73-104 semantics across GPUs).  With NCCL that
// is an all-gather of W * Q * k records into a staging buffer on every rank followed by the merge kernel.
// Here every rank's buffers are exported with CUDA IPC and mapped by its peers, and
//
//   * rank r merges only ITS SLICE of the queries (Q / W of them): its merge kernel reads the W lists
//     of those queries straight from the owners' HBM over NVLink / NVSwitch (16-byte loads, the W * k
//     records of a query in flight together) while it ranks them -- each record crosses the fabric
//     once, an all-gather moves it W - 1 times;
//   * the merged rows are stored straight into every rank's result buffer (posted 16-byte writes);
//   * no staging copy, no collective, no host round trip on the step path.
//
// Synchronisation is two step counters per rank in the exported buffer.  `ready`: after its local
// top-k a rank publishes the step (system-scope release); a merge kernel waits (acquire) until all W
// ranks are ready before it reads.  `done`: after its merge kernel a rank publishes the step again;
// a rank's result is complete once all W ranks are done.  Lists and results are double-buffered by
// step parity: slot s%2 is rewritten at step s+2, which a rank can only reach after every peer has
// published step s+1 -- and each peer publishes s+1 after its own step s has consumed everything it
// reads.  Every wait is bounded (a missing peer sets an error word the host can read, the kernel
// returns): a dead rank costs a failed step, not a hung GPU.
#include <algorithm>
#include <vector>

#include "nwk_common.cuh"

namespace nwk {

constexpr int XCHG_MAX_WORLD = 16;
constexpr int XCHG_THREADS = 128;             // 4 warps = 4 queries per CTA (merge_rank layout)
constexpr unsigned long long XCHG_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;

struct XchgPeers {
  const unsigned char* buf[XCHG_MAX_WORLD];   // peer r's exported buffer (own buffer for r == rank)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ size_t align_up_dev(size_t x, size_t a) { return (x + a - 1) / a * a; }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// the rank's lists for `step` are complete (stream order): make them visible system-wide, then publish
__global__ void xchg_publish_kernel(unsigned int* flag, unsigned int step) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
}

// bounded wait until every rank's counter at `off` has reached `step`: thread r watches rank r
__device__ __forceinline__ bool wait_all(const XchgPeers& peers, int world, size_t off, unsigned int step, int* err) {
  bool ok = true;
  if (threadIdx.x < world) {
    const unsigned int* f = reinterpret_cast<const unsigned int*>(peers.buf[threadIdx.x] + off);
    const unsigned long long t0 = globaltimer_ns();
    // (steps are compared modulo 2^32 with a signed difference: the counter may wrap)
    while ((int)(ld_acquire_sys(f) - step) < 0) {
      if (globaltimer_ns() - t0 > XCHG_TIMEOUT_NS) { ok = false; break; }
      __nanosleep(200);
    }
    if (!ok) atomicExch(err, 1 + (int)threadIdx.x);
  }
  // (the barrier orders every thread's later reads after the acquiring loads)
  return __syncthreads_and(ok);
}

struct XchgGeom {
  size_t list_off[2], res_off[2];     // byte offsets of the two list slots / result slots in a rank's buffer
  size_t idx_off;                     // ids follow the distances inside a slot at this offset
  size_t ready_off, done_off;         // the two step counters
};

// gather + merge + broadcast for this rank's slice of the queries [q_lo, q_hi): one warp per query.
// List r of a query is read from peer r; the merged row is written to every peer's result slot.
// Order and padding rules are those of merge_rank_kernel (nwk_scan.cu): (distance, id, list, position);
// a negative id marks padding; lists sorted by distance are ranked by binary search.
// k_in % 4 == 0 and 16-byte aligned slots let a lane move 16 bytes per access.
__global__ void __launch_bounds__(XCHG_THREADS)
xchg_merge_kernel(const XchgPeers peers, const XchgGeom g, int world, int slot, unsigned int step, long long q_lo,
                  long long q_hi, long long nq, int k_in, int k_out, int vec, int* __restrict__ err,
                  unsigned int* __restrict__ cta_counter, unsigned int* __restrict__ done_flag) {
  extern __shared__ __align__(16) unsigned char xm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!wait_all(peers, world, g.ready_off, step, err)) return;

  const int n = world * k_in;
  const int n_out = (k_out + 3) & ~3;                          // staged output row, padded to 16 bytes
  long long* s_id = reinterpret_cast<long long*>(xm_smem) + (size_t)warp * (n + n_out);
  long long* s_oid = s_id + n;
  float* s_d = reinterpret_cast<float*>(xm_smem + (size_t)(XCHG_THREADS / 32) * (n + n_out) * 8) + (size_t)warp * (n + n_out);
  float* s_od = s_d + n;
  const long long q = q_lo + (long long)blockIdx.x * (XCHG_THREADS / 32) + warp;
  const float INF = __int_as_float(0x7f800000);
  const long long PAD = 0x7fffffffffffffffLL;
  if (q < q_hi) {
  // ---- gather: the W lists of the query, from their owners (.cg: peer lines never sit in this SM's L1) ----
  if (vec) {
    const int f4 = k_in >> 2, l2 = k_in >> 1;                  // float4 / longlong2 per list
    for (int e = lane; e < world * f4; e += 32) {
      const int r = e / f4, j = e - r * f4;
      const float4 v = __ldcg(reinterpret_cast<const float4*>(peers.buf[r] + g.list_off[slot]) + (size_t)q * f4 + j);
      *reinterpret_cast<float4*>(s_d + r * k_in + 4 * j) = v;
    }
    for (int e = lane; e < world * l2; e += 32) {
      const int r = e / l2, j = e - r * l2;
      const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(peers.buf[r] + g.list_off[slot] + g.idx_off) + (size_t)q * l2 + j);
      *reinterpret_cast<longlong2*>(s_id + r * k_in + 2 * j) = v;
    }
  } else {
    for (int e = lane; e < n; e += 32) {
      const int r = e / k_in, j = e - r * k_in;
      const unsigned char* sl = peers.buf[r] + g.list_off[slot];
      s_d[e] = __ldcg(reinterpret_cast<const float*>(sl) + (size_t)q * k_in + j);
      s_id[e] = __ldcg(reinterpret_cast<const long long*>(sl + g.idx_off) + (size_t)q * k_in + j);
    }
  }
  __syncwarp();
  for (int e = lane; e < n; e += 32) {
    float d = s_d[e];
    long long id = s_id[e];
    if (!(d == d)) d = INF;
    if (id < 0) { d = INF; id = PAD; }
    s_d[e] = d;
    s_id[e] = id;
  }
  __syncwarp();
  bool sorted = true;
  for (int e = lane; e < n; e += 32)
    if ((e % k_in) != 0 && s_d[e - 1] > s_d[e]) sorted = false;
  sorted = __all_sync(0xffffffffu, sorted);
  // ---- rank: (distance, id, position) is a strict total order, so every output slot has one writer ----
  for (int e = lane; e < n; e += 32) {
    const float d = s_d[e];
    const long long id = s_id[e];
    int rank = 0;
    if (sorted) {
      for (int b = 0; b < world; ++b) {
        const float* l = s_d + b * k_in;
        int lo = 0, hi = k_in;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (l[mid] < d) lo = mid + 1; else hi = mid;
        }
        rank += lo;
        for (int pos = lo; pos < k_in && l[pos] == d; ++pos) {
          const int f = b * k_in + pos;
          const long long fid = s_id[f];
          if (fid < id || (fid == id && f < e)) ++rank;
        }
      }
    } else {
      for (int f = 0; f < n; ++f) {
        const float fd = s_d[f];
        const long long fid = s_id[f];
        if (fd < d || (fd == d && (fid < id || (fid == id && f < e)))) ++rank;
      }
    }
    if (rank < k_out) { s_od[rank] = d; s_oid[rank] = (id == PAD) ? -1 : id; }
  }
  __syncwarp();
  // ---- broadcast: the merged row into every rank's result slot ----
  const size_t res_idx_off = align_up_dev((size_t)nq * k_out * 4, 256);
  if (vec && (k_out & 3) == 0) {
    const int f4 = k_out >> 2, l2 = k_out >> 1;
    for (int e = lane; e < world * f4; e += 32) {
      const int r = e / f4, j = e - r * f4;
      reinterpret_cast<float4*>(const_cast<unsigned char*>(peers.buf[r]) + g.res_off[slot])[(size_t)q * f4 + j] =
          *reinterpret_cast<const float4*>(s_od + 4 * j);
    }
    for (int e = lane; e < world * l2; e += 32) {
      const int r = e / l2, j = e - r * l2;
      reinterpret_cast<longlong2*>(const_cast<unsigned char*>(peers.buf[r]) + g.res_off[slot] + res_idx_off)[(size_t)q * l2 + j] =
          *reinterpret_cast<const longlong2*>(s_oid + 2 * j);
    }
  } else {
    for (int e = lane; e < world * k_out; e += 32) {
      const int r = e / k_out, j = e - r * k_out;
      unsigned char* rs = const_cast<unsigned char*>(peers.buf[r]) + g.res_off[slot];
      reinterpret_cast<float*>(rs)[(size_t)q * k_out + j] = s_od[j];
      reinterpret_cast<long long*>(rs + res_idx_off)[(size_t)q * k_out + j] = s_oid[j];
    }
  }
  }
  // ---- done: every writer fences its peer stores system-wide; the last CTA publishes the step ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(cta_counter, 1u);
    if (prev == gridDim.x - 1) {
      *cta_counter = 0u;                       // (the next launch is ordered after this kernel on the stream)
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(done_flag), "r"(step) : "memory");
    }
  }
}

// the result slot of this rank is complete once every rank has published `done` for the step
__global__ void xchg_wait_kernel(const XchgPeers peers, int world, size_t done_off, unsigned int step, int* err) {
  wait_all(peers, world, done_off, step, err);
}

}  // namespace nwk

using namespace nwk;

struct nwk_xchg {
  int rank, world, k, device;
  long long nq;
  XchgGeom g;
  size_t slot_bytes, total_bytes;
  unsigned char* local = nullptr;              // [2 list slots][2 result slots][counters]
  unsigned char* peer[XCHG_MAX_WORLD] = {};
  bool opened[XCHG_MAX_WORLD] = {};
  unsigned int step = 0;                       // last published step
  int* err_host = nullptr;                     // mapped pinned word the kernels set on a timeout
  int* err_dev = nullptr;
  int smem_optin = 48 * 1024;
};

extern "C" {

int nwk_xchg_create(int32_t rank, int32_t world, int64_t nq, int32_t k, nwk_xchg** out) {
  if (!out || world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world || nq <= 0 || k <= 0 || k > NWK_MAX_K)
    return fail(NWK_ERR_ARG, "bad exchange shape (world <= %d)", XCHG_MAX_WORLD);
  if ((long long)world * k > 2048) return fail(NWK_ERR_ARG, "world * k = %lld exceeds 2048 records per query", (long long)world * k);
  NWK_TRY(device_check());
  nwk_xchg* x = new nwk_xchg();
  x->rank = rank; x->world = world; x->k = k; x->nq = nq;
  cudaGetDevice(&x->device);
  x->g.idx_off = align_up((size_t)nq * k * 4, 256);
  x->slot_bytes = align_up(x->g.idx_off + (size_t)nq * k * 8, 256);
  for (int i = 0; i < 2; ++i) {
    x->g.list_off[i] = (size_t)i * x->slot_bytes;
    x->g.res_off[i] = (size_t)(2 + i) * x->slot_bytes;
  }
  x->g.ready_off = 4 * x->slot_bytes;
  x->g.done_off = x->g.ready_off + 128;
  x->total_bytes = x->g.ready_off + 512;       // ready, done, and (local only) the CTA counter of the merge kernel at +256
  cudaError_t e = cudaMalloc(&x->local, x->total_bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->total_bytes);
  if (e == cudaSuccess) e = cudaHostAlloc(&x->err_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *x->err_host = 0; e = cudaHostGetDevicePointer(&x->err_dev, x->err_host, 0); }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    nwk_xchg_destroy(x);
    return fail(NWK_ERR_CUDA, "exchange buffer allocation failed: %s", cudaGetErrorString(e));
  }
  x->peer[rank] = x->local;
  *out = x;
  return NWK_OK;
}

size_t nwk_xchg_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int nwk_xchg_local_handle(nwk_xchg* x, void* handle_out) {
  if (!x || !handle_out) return fail(NWK_ERR_ARG, "null pointer argument");
  cudaIpcMemHandle_t h;
  NWK_CUDA(cudaIpcGetMemHandle(&h, x->local));
  memcpy(handle_out, &h, sizeof(h));
  return NWK_OK;
}

// handles: world x nwk_xchg_handle_bytes(), entry r exported by rank r (entry `rank` is ignored)
int nwk_xchg_open_peers(nwk_xchg* x, const void* handles) {
  if (!x || !handles) return fail(NWK_ERR_ARG, "null pointer argument");
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank || x->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(NWK_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
    x->peer[r] = static_cast<unsigned char*>(p);
    x->opened[r] = true;
  }
  return NWK_OK;
}

// where the local top-k of the NEXT step is to be written (device pointers into the exported buffer)
int nwk_xchg_next_slot(nwk_xchg* x, float** dist, int64_t** idx) {
  if (!x || !dist || !idx) return fail(NWK_ERR_ARG, "null pointer argument");
  unsigned char* slot = x->local + x->g.list_off[(x->step + 1) & 1u];
  *dist = reinterpret_cast<float*>(slot);
  *idx = reinterpret_cast<int64_t*>(slot + x->g.idx_off);
  return NWK_OK;
}

// publish the lists written to the slot of nwk_xchg_next_slot, merge this rank's slice of the queries out of
// the peers' memory into every rank's result slot, wait for the peers' slices, copy out.  Enqueued on `stream`.
int nwk_xchg_merge(nwk_xchg* x, int32_t k_out, float* out_dist, int64_t* out_idx, void* stream) {
  if (!x || !out_dist || !out_idx) return fail(NWK_ERR_ARG, "null pointer argument");
  if (k_out <= 0 || k_out > x->k) return fail(NWK_ERR_ARG, "k_out must be in [1, k]");
  if (*reinterpret_cast<volatile int*>(x->err_host) != 0)
    return fail(NWK_ERR_INTERNAL, "peer exchange timed out waiting for rank %d", *x->err_host - 1);
  for (int r = 0; r < x->world; ++r)
    if (!x->peer[r]) return fail(NWK_ERR_ARG, "peer %d not opened (nwk_xchg_open_peers)", r);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  x->step += 1;
  const int slot = (int)(x->step & 1u);
  xchg_publish_kernel<<<1, 1, 0, s>>>(reinterpret_cast<unsigned int*>(x->local + x->g.ready_off), x->step);
  NWK_LAUNCH_CHECK();
  XchgPeers peers{};
  for (int r = 0; r < x->world; ++r) peers.buf[r] = x->peer[r];
  // this rank's slice of the queries
  const long long base_n = x->nq / x->world, extra = x->nq % x->world;
  const long long q_lo = x->rank * base_n + std::min<long long>(x->rank, extra);
  const long long q_hi = q_lo + base_n + (x->rank < extra ? 1 : 0);
  const int n = x->world * x->k, n_out = (k_out + 3) & ~3;
  const int smem = (XCHG_THREADS / 32) * (n + n_out) * 12;
  if (smem > x->smem_optin) {
    NWK_CUDA(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    x->smem_optin = smem;
  }
  const long long blocks = std::max<long long>(1, (q_hi - q_lo + XCHG_THREADS / 32 - 1) / (XCHG_THREADS / 32));
  const int vec = (x->k % 4 == 0) ? 1 : 0;     // 16-byte accesses need rows that start on 16-byte boundaries
  xchg_merge_kernel<<<(unsigned)blocks, XCHG_THREADS, smem, s>>>(peers, x->g, x->world, slot, x->step, q_lo, q_hi, x->nq,
                                                                x->k, k_out, vec, x->err_dev,
                                                                reinterpret_cast<unsigned int*>(x->local + x->g.ready_off + 256),
                                                                reinterpret_cast<unsigned int*>(x->local + x->g.done_off));
  NWK_LAUNCH_CHECK();
  xchg_wait_kernel<<<1, 32, 0, s>>>(peers, x->world, x->g.done_off, x->step, x->err_dev);
  NWK_LAUNCH_CHECK();
  const unsigned char* res = x->local + x->g.res_off[slot];
  const size_t res_idx_off = align_up((size_t)x->nq * k_out * 4, 256);
  NWK_CUDA(cudaMemcpyAsync(out_dist, res, (size_t)x->nq * k_out * 4, cudaMemcpyDeviceToDevice, s));
  NWK_CUDA(cudaMemcpyAsync(out_idx, res + res_idx_off, (size_t)x->nq * k_out * 8, cudaMemcpyDeviceToDevice, s));
  return NWK_OK;
}

// 0 = healthy; otherwise the call fails naming the rank a kernel gave up waiting for
int nwk_xchg_check(nwk_xchg* x) {
  if (!x) return fail(NWK_ERR_ARG, "null pointer argument");
  const int e = *reinterpret_cast<volatile int*>(x->err_host);
  if (e != 0) return fail(NWK_ERR_INTERNAL, "peer exchange timed out waiting for rank %d", e - 1);
  return NWK_OK;
}

int nwk_xchg_destroy(nwk_xchg* x) {
  if (!x) return NWK_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (x->opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->local) cudaFree(x->local);
  if (x->err_host) cudaFreeHost(x->err_host);
  delete x;
  return NWK_OK;
}

}  // extern "C"
