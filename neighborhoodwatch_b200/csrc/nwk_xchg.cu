// nwk_xchg.cu -- the cross-GPU reduction as ONE kernel over peer memory: gather + k-way merge.
//
// Multi-GPU step (one process per GPU): every rank produces its exact local top-k [Q, k] with global
// ids, then the W lists of a query are merged (merge.py:73-104 semantics across GPUs).  With NCCL that
// is an all-gather into a staging buffer followed by the merge kernel.  Here the lists never move
// twice: every rank's result buffer is exported with CUDA IPC, mapped by its peers, and the merge
// kernel reads the W lists of its queries straight from the owners' HBM over NVLink / NVSwitch while
// it ranks them -- no collective call, no staging copy, no host round trip.
//
// Synchronisation is a step counter per rank in the exported buffer: after its local top-k a rank
// publishes `step` (system-scope release), the merge kernel of every rank waits until all W counters
// have reached `step` (acquire) before it reads.  Results are double-buffered by step parity: a rank
// can only overwrite slot s%2 at step s+2, after its own merge of step s+1 has seen every peer
// publish s+1 -- which each peer does after finishing its merge of step s, the last reader of slot
// s%2.  Every wait is bounded (a missing peer sets an error word the host can read, the kernel
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

// gather + merge: one warp per query; list r of the query is read from peer r's slot.
// Order and padding rules are those of merge_rank_kernel (nwk_scan.cu): (distance, id, list, position);
// a negative id marks padding; lists sorted by distance are ranked by binary search.
__global__ void __launch_bounds__(XCHG_THREADS)
xchg_merge_kernel(const XchgPeers peers, int world, size_t slot_off, size_t idx_off, size_t flag_off, unsigned int step,
                  long long nq, int k_in, int k_out, float* __restrict__ out_dist, long long* __restrict__ out_idx,
                  int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char xm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- wait until every rank has published this step (bounded): thread r watches rank r ----
  bool ok = true;
  if (threadIdx.x < world) {
    const unsigned int* f = reinterpret_cast<const unsigned int*>(peers.buf[threadIdx.x] + flag_off);
    const unsigned long long t0 = globaltimer_ns();
    // (steps are compared modulo 2^32 with a signed difference: the counter may wrap)
    while ((int)(ld_acquire_sys(f) - step) < 0) {
      if (globaltimer_ns() - t0 > XCHG_TIMEOUT_NS) { ok = false; break; }
      __nanosleep(200);
    }
    if (!ok) atomicExch(err, 1 + (int)threadIdx.x);
  }
  // (the barrier orders every thread's reads of the peers' lists after the acquiring loads)
  if (!__syncthreads_and(ok)) return;

  const int n = world * k_in;
  long long* s_id = reinterpret_cast<long long*>(xm_smem) + (size_t)warp * n;
  float* s_d = reinterpret_cast<float*>(xm_smem + (size_t)(XCHG_THREADS / 32) * n * 8) + (size_t)warp * n;
  const long long q = (long long)blockIdx.x * (XCHG_THREADS / 32) + warp;
  if (q >= nq) return;
  const float INF = __int_as_float(0x7f800000);
  const long long PAD = 0x7fffffffffffffffLL;
  for (int e = lane; e < n; e += 32) {
    const int r = e / k_in, j = e - r * k_in;
    const unsigned char* slot = peers.buf[r] + slot_off;
    // (.cg: peer lines are never served from this SM's L1)
    float d = __ldcg(reinterpret_cast<const float*>(slot) + (size_t)q * k_in + j);
    long long id = __ldcg(reinterpret_cast<const long long*>(slot + idx_off) + (size_t)q * k_in + j);
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
    if (rank < k_out) { out_dist[q * k_out + rank] = d; out_idx[q * k_out + rank] = (id == PAD) ? -1 : id; }
  }
}

}  // namespace nwk

using namespace nwk;

struct nwk_xchg {
  int rank, world, k, device;
  long long nq;
  size_t slot_bytes, idx_off, flag_off, total_bytes;
  unsigned char* local = nullptr;              // [2 slots][dist | idx] + flags
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
  x->idx_off = align_up((size_t)nq * k * 4, 256);
  x->slot_bytes = align_up(x->idx_off + (size_t)nq * k * 8, 256);
  x->flag_off = 2 * x->slot_bytes;
  x->total_bytes = x->flag_off + 256;
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
  unsigned char* slot = x->local + (size_t)((x->step + 1) & 1u) * x->slot_bytes;
  *dist = reinterpret_cast<float*>(slot);
  *idx = reinterpret_cast<int64_t*>(slot + x->idx_off);
  return NWK_OK;
}

// publish the lists written to the slot of nwk_xchg_next_slot, wait for every peer's, merge.  Enqueued on `stream`.
int nwk_xchg_merge(nwk_xchg* x, int32_t k_out, float* out_dist, int64_t* out_idx, void* stream) {
  if (!x || !out_dist || !out_idx) return fail(NWK_ERR_ARG, "null pointer argument");
  if (k_out <= 0 || k_out > x->world * x->k) return fail(NWK_ERR_ARG, "bad k_out");
  if (*reinterpret_cast<volatile int*>(x->err_host) != 0)
    return fail(NWK_ERR_INTERNAL, "peer exchange timed out waiting for rank %d", *x->err_host - 1);
  for (int r = 0; r < x->world; ++r)
    if (!x->peer[r]) return fail(NWK_ERR_ARG, "peer %d not opened (nwk_xchg_open_peers)", r);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  x->step += 1;
  const size_t slot_off = (size_t)(x->step & 1u) * x->slot_bytes;
  xchg_publish_kernel<<<1, 1, 0, s>>>(reinterpret_cast<unsigned int*>(x->local + x->flag_off), x->step);
  NWK_LAUNCH_CHECK();
  XchgPeers peers{};
  for (int r = 0; r < x->world; ++r) peers.buf[r] = x->peer[r];
  const int n = x->world * x->k;
  const int smem = (XCHG_THREADS / 32) * n * 12;
  if (smem > x->smem_optin) {
    NWK_CUDA(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    x->smem_optin = smem;
  }
  const long long blocks = (x->nq + XCHG_THREADS / 32 - 1) / (XCHG_THREADS / 32);
  xchg_merge_kernel<<<(unsigned)blocks, XCHG_THREADS, smem, s>>>(peers, x->world, slot_off, x->idx_off, x->flag_off, x->step,
                                                                x->nq, x->k, k_out, out_dist, (long long*)out_idx, x->err_dev);
  NWK_LAUNCH_CHECK();
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
