// nwk_api.cu -- the extern "C" surface declared in include/nwk.h.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "nwk_common.cuh"

namespace nwk {

// implemented in the other translation units
int exact_scan(const float* base, long long nb, const float* query, int d, int k, int metric,
               long long id_offset, const int* qsel, long long nsel, float* out_dist,
               long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s);
size_t exact_scan_workspace_bytes(long long nb, long long nsel, int k);
size_t merge_workspace_bytes(int P, long long nq, int k_in, int k_out);
int merge_topk(int P, const float* dist, const long long* idx, long long nq, int k_in, int k_out,
               float* out_dist, long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s);
int merge_topk_strided(int P, const float* dist, long long dist_stride, const long long* idx, long long idx_stride,
                       long long nq, int k_in, int k_out, float* out_dist, long long* out_idx, void* ws,
                       size_t ws_bytes, cudaStream_t s);
int merge_topk_ptrs(int P, const float* const* dist, const long long* const* idx, long long nq, int k_in, int k_out,
                    float* out_dist, long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s);
bool tc_path_supported(long long nb, long long nq, int d, int k);
size_t tc_workspace_bytes(long long nb, long long nq, int d, int k);
int tc_knn(const float* base, long long nb, const float* query, long long nq, int d, int k, int metric,
           long long id_offset, float* out_dist, long long* out_idx, void* ws, size_t ws_bytes,
           unsigned int flags, nwk_knn_stats* stats, cudaStream_t s, const float* bound, long long bound_stride);
int tc_debug_tile(const float* base, long long nb, const float* query, long long nq, int d, long long b0,
                  long long q0, float* out, int use_f16, cudaStream_t s);
int synth_fill(float* out, long long rows, int d, unsigned long long seed, long long row0, int normalize,
               cudaStream_t s);
size_t maxsim_workspace_bytes(long long nq, int Tq, long long nd, int Td, int d, int k);
int maxsim_topk(const float* qtok, long long nq, int Tq, const float* docs, long long nd, int Td, const int* doc_len, int d,
                int k, long long id_offset, float* out_score, long long* out_idx, void* ws, size_t ws_bytes,
                nwk_knn_stats* stats, cudaStream_t s);

int rows_from_columns(const float* src, long long ld, int d, long long rows, float* dst, cudaStream_t s);
int validate_pairs(const float* base, long long nb, const float* query, long long nq, int d, const long long* idx,
                   const float* dist, int k, double rtol, double atol, long long* out_host, cudaStream_t s);

std::atomic<long long> g_launches{0};

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

static DevState g_dev[NWK_MAX_DEVICES];
static std::mutex g_dev_mu;

int device_check() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return fail(NWK_ERR_DEVICE, "no CUDA device available: %s (this library has no CPU fallback)",
                cudaGetErrorString(e));
  if (dev < 0 || dev >= NWK_MAX_DEVICES) return fail(NWK_ERR_DEVICE, "device ordinal %d out of range", dev);
  DevState& st = g_dev[dev];
  if (st.checked) return NWK_OK;
  std::lock_guard<std::mutex> lock(g_dev_mu);
  if (st.checked) return NWK_OK;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return fail(NWK_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(NWK_ERR_DEVICE, "device %d (%s) is sm_%d%d; libnwk is built for sm_100a only", dev, prop.name,
                prop.major, prop.minor);
  st.sms = prop.multiProcessorCount;
  for (int i = 0; i < NWK_SIDE_STREAMS; ++i) {
    NWK_CUDA(cudaStreamCreateWithFlags(&st.side[i], cudaStreamNonBlocking));
    NWK_CUDA(cudaEventCreateWithFlags(&st.ev_join[i], cudaEventDisableTiming));
  }
  NWK_CUDA(cudaEventCreateWithFlags(&st.ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 16; ++i) NWK_CUDA(cudaEventCreateWithFlags(&st.ev_chunk[i], cudaEventDisableTiming));
  st.checked = 1;
  return NWK_OK;
}

// (callers have passed device_check() on this device)
DevState* dev_state() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= NWK_MAX_DEVICES) dev = 0;
  return &g_dev[dev];
}

int sm_count() {
  const int n = dev_state()->sms;
  return n > 0 ? n : 148;
}

static bool want_tc(long long nb, long long nq, int d, int k, unsigned int flags) {
  if (flags & NWK_FLAG_FORCE_EXACT_SCAN) return false;
  if (getenv("NWK_FORCE_EXACT_SCAN")) return false;
  return tc_path_supported(nb, nq, d, k);
}

}  // namespace nwk

using namespace nwk;

__global__ void state_init_kernel(float* d, long long* i, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    d[t] = __int_as_float(0x7f800000);
    i[t] = -1;
  }
}

// Per-device cache of the host path's staging buffers, streams and running state, so repeated
// calls pay no cudaMalloc / cudaFree / device-wide synchronisation (include/nwk.h: "no hidden
// allocations per call" once the buffers have grown to the shapes in use).
namespace {
struct HostCtx {
  std::mutex mu;
  float* slab[2] = {nullptr, nullptr};
  size_t slab_bytes[2] = {0, 0};
  float* query = nullptr;
  size_t query_bytes = 0;
  char* ws = nullptr;
  size_t ws_bytes = 0;
  float* part_d = nullptr;
  size_t part_d_bytes = 0;
  long long* part_i = nullptr;
  size_t part_i_bytes = 0;
  nwk_topk_state* state = nullptr;      // running result of multi-slab calls, reused while (nq, k) stay the same
  cudaStream_t copy_stream = nullptr, comp_stream = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  bool init = false;
};
HostCtx g_host[NWK_MAX_DEVICES];

template <typename T>
int grow(T** p, size_t* have, size_t want) {
  if (*have >= want && *p) return NWK_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) {
    *p = nullptr;
    return fail(NWK_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  *have = want;
  return NWK_OK;
}
}  // namespace

extern "C" {

const char* nwk_last_error(void) { return last_error_ref().c_str(); }
int nwk_abi_version(void) { return NWK_ABI_VERSION; }
int64_t nwk_launch_count(void) { return (int64_t)g_launches.load(); }

size_t nwk_knn_workspace_bytes(int64_t nb, int64_t nq, int32_t d, int32_t k, int32_t metric, uint32_t flags) {
  (void)metric;
  if (nb <= 0 || nq <= 0 || d <= 0 || k <= 0) return 256;
  if (want_tc(nb, nq, d, k, flags)) return tc_workspace_bytes(nb, nq, d, k);
  return exact_scan_workspace_bytes(nb, nq, k) + 256;
}

// nwk_knn_f32, optionally bounded by the caller's current k-th distances (bound[q * bound_stride], device)
static int knn_impl(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d, int32_t k,
                    int32_t metric, int64_t id_offset, float* out_dist, int64_t* out_idx, void* workspace,
                    size_t ws_bytes, uint32_t flags, nwk_knn_stats* stats, void* stream, const float* bound,
                    long long bound_stride) {
  if (!base || !query || !out_dist || !out_idx) return fail(NWK_ERR_ARG, "null pointer argument");
  flags &= 0x7fffffffu;       // (the top bit is the library's own: a call that is a second-chance pass)
  if (nb <= 0 || nq < 0 || d <= 0) return fail(NWK_ERR_ARG, "bad shape nb=%lld nq=%lld d=%d", (long long)nb, (long long)nq, d);
  if (k <= 0 || k > NWK_MAX_K) return fail(NWK_ERR_ARG, "k=%d outside [1, %d]", k, NWK_MAX_K);
  if (k > nb) return fail(NWK_ERR_ARG, "k=%d exceeds dataset rows %lld", k, (long long)nb);
  if (metric < 0 || metric > 2) return fail(NWK_ERR_ARG, "unknown metric %d", metric);
  if (nb >= (1ll << 31)) return fail(NWK_ERR_ARG, "nb=%lld: at most 2^31-1 base rows per call", (long long)nb);
  NWK_TRY(device_check());
  if (stats) memset(stats, 0, sizeof(*stats));
  if (nq == 0) return NWK_OK;
  if (ws_bytes < nwk_knn_workspace_bytes(nb, nq, d, k, metric, flags) || !workspace)
    return fail(NWK_ERR_WORKSPACE, "workspace of %zu bytes is smaller than nwk_knn_workspace_bytes() = %zu",
                ws_bytes, nwk_knn_workspace_bytes(nb, nq, d, k, metric, flags));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // (the tensor-core path reads rows as 16-byte vectors and through TMA; base or query pointers that
  // are not 16-byte aligned take the fp64 scan, in rounds that fit the same workspace)
  const bool aligned = ((reinterpret_cast<uintptr_t>(base) | reinterpret_cast<uintptr_t>(query)) & 15u) == 0;
  if (want_tc(nb, nq, d, k, flags) && aligned)
    return tc_knn(base, nb, query, nq, d, k, metric, id_offset, out_dist, (long long*)out_idx, workspace,
                  ws_bytes, flags, stats, s, bound, bound_stride);
  // The fp64 scan ignores the bound: it returns the slab's full top-k, a superset of what the caller
  // needs.  The workspace may have been sized for the tensor-core path (which provisions the scan for
  // 64 queries): the scan then walks the queries in rounds that fit, as the fallback does.
  const long long l0 = g_launches.load();
  long long per = nq;
  while (per > 8 && exact_scan_workspace_bytes(nb, per, k) > ws_bytes) per = (per + 1) / 2;
  if (exact_scan_workspace_bytes(nb, per, k) > ws_bytes)
    return fail(NWK_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one round of the exact scan", ws_bytes);
  for (long long q0 = 0; q0 < nq; q0 += per) {
    const long long n = std::min<long long>(per, nq - q0);
    NWK_TRY(exact_scan(base, nb, query + (size_t)q0 * d, d, k, metric, id_offset, nullptr, n, out_dist + (size_t)q0 * k,
                       (long long*)out_idx + (size_t)q0 * k, workspace, ws_bytes, s));
  }
  NWK_CUDA(cudaStreamSynchronize(s));
  if (stats) {
    stats->path = 0;
    stats->gpu_launches = (int)(g_launches.load() - l0);
  }
  return NWK_OK;
}

int nwk_knn_f32(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d, int32_t k,
                int32_t metric, int64_t id_offset, float* out_dist, int64_t* out_idx, void* workspace,
                size_t ws_bytes, uint32_t flags, nwk_knn_stats* stats, void* stream) {
  return knn_impl(base, nb, query, nq, d, k, metric, id_offset, out_dist, out_idx, workspace, ws_bytes, flags, stats,
                  stream, nullptr, 0);
}

size_t nwk_merge_workspace_bytes(int32_t P, int64_t nq, int32_t k_in, int32_t k_out) {
  if (P <= 0 || nq <= 0 || k_in <= 0 || k_out <= 0) return 256;
  return merge_workspace_bytes(P, nq, k_in, k_out);
}

static int merge_args_check(int32_t P, const void* dist, const void* idx, int64_t nq, int32_t k_in, int32_t k_out,
                            const void* out_dist, const void* out_idx) {
  if (!dist || !idx || !out_dist || !out_idx) return fail(NWK_ERR_ARG, "null pointer argument");
  if (P <= 0 || nq < 0 || k_in <= 0 || k_out <= 0) return fail(NWK_ERR_ARG, "bad merge shape");
  if (k_out > NWK_MERGE_MAX_K) return fail(NWK_ERR_ARG, "k_out=%d exceeds %d", k_out, NWK_MERGE_MAX_K);
  if ((long long)k_out > (long long)P * k_in)
    return fail(NWK_ERR_ARG, "k_out=%d exceeds the %lld candidates per query", k_out, (long long)P * k_in);
  return device_check();
}

int nwk_merge_topk(int32_t P, const float* dist, const int64_t* idx, int64_t nq, int32_t k_in, int32_t k_out,
                   float* out_dist, int64_t* out_idx, void* workspace, size_t ws_bytes, void* stream) {
  NWK_TRY(merge_args_check(P, dist, idx, nq, k_in, k_out, out_dist, out_idx));
  return merge_topk(P, dist, (const long long*)idx, nq, k_in, k_out, out_dist, (long long*)out_idx, workspace,
                    ws_bytes, static_cast<cudaStream_t>(stream));
}

int nwk_merge_topk_strided(int32_t P, const float* dist, int64_t dist_stride, const int64_t* idx, int64_t idx_stride,
                           int64_t nq, int32_t k_in, int32_t k_out, float* out_dist, int64_t* out_idx, void* workspace,
                           size_t ws_bytes, void* stream) {
  NWK_TRY(merge_args_check(P, dist, idx, nq, k_in, k_out, out_dist, out_idx));
  return merge_topk_strided(P, dist, dist_stride, (const long long*)idx, idx_stride, nq, k_in, k_out, out_dist,
                            (long long*)out_idx, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}

int nwk_merge_topk_ptrs(int32_t P, const float* const* dist_ptrs, const int64_t* const* idx_ptrs, int64_t nq,
                        int32_t k_in, int32_t k_out, float* out_dist, int64_t* out_idx, void* workspace,
                        size_t ws_bytes, void* stream) {
  NWK_TRY(merge_args_check(P, dist_ptrs, idx_ptrs, nq, k_in, k_out, out_dist, out_idx));
  for (int p = 0; p < P && p < 16; ++p)
    if (!dist_ptrs[p] || !idx_ptrs[p]) return fail(NWK_ERR_ARG, "null list pointer %d", p);
  return merge_topk_ptrs(P, dist_ptrs, reinterpret_cast<const long long* const*>(idx_ptrs), nq, k_in, k_out, out_dist,
                         (long long*)out_idx, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}

// ---- running top-k state --------------------------------------------------------------
struct nwk_topk_state {
  long long nq;
  int k;
  int device;
  float* pack_d;       // [2][nq][k]: slot 0 = current result, slot 1 = incoming partial
  long long* pack_i;
  float* tmp_d;        // [nq][k]
  long long* tmp_i;
  void* ws;
  size_t ws_bytes;
};

int nwk_topk_state_reset(nwk_topk_state* st, void* stream) {
  if (!st) return fail(NWK_ERR_ARG, "null pointer argument");
  state_init_kernel<<<256, 256, 0, static_cast<cudaStream_t>(stream)>>>(st->pack_d, st->pack_i, (long long)st->nq * st->k);
  NWK_LAUNCH_CHECK();
  return NWK_OK;
}

int nwk_topk_state_create(int64_t nq, int32_t k, nwk_topk_state** out) {
  if (!out || nq <= 0 || k <= 0 || k > NWK_MAX_K) return fail(NWK_ERR_ARG, "bad topk state shape");
  NWK_TRY(device_check());
  nwk_topk_state* st = new nwk_topk_state();
  st->nq = nq;
  st->k = k;
  cudaGetDevice(&st->device);
  const size_t n = (size_t)nq * k;
  st->ws_bytes = merge_workspace_bytes(2, nq, k, k);
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&st->pack_d, 2 * n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&st->pack_i, 2 * n * sizeof(long long));
  if (e == cudaSuccess) e = cudaMalloc(&st->tmp_d, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&st->tmp_i, n * sizeof(long long));
  if (e == cudaSuccess) e = cudaMalloc(&st->ws, st->ws_bytes);
  if (e != cudaSuccess) {
    nwk_topk_state_destroy(st);
    return fail(NWK_ERR_CUDA, "topk state allocation failed: %s", cudaGetErrorString(e));
  }
  // (the legacy default stream orders this initialisation before any later work of blocking streams;
  //  callers on non-blocking streams are covered by the synchronisation)
  int rc = nwk_topk_state_reset(st, nullptr);
  if (rc == NWK_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = fail(NWK_ERR_CUDA, "topk state initialisation failed");
  if (rc != NWK_OK) { nwk_topk_state_destroy(st); return rc; }
  *out = st;
  return NWK_OK;
}

// the running result moves from tmp back into slot 0 (two device copies, enqueued)
static int state_commit(nwk_topk_state* st, cudaStream_t s) {
  const size_t n = (size_t)st->nq * st->k;
  NWK_CUDA(cudaMemcpyAsync(st->pack_d, st->tmp_d, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  NWK_CUDA(cudaMemcpyAsync(st->pack_i, st->tmp_i, n * sizeof(long long), cudaMemcpyDeviceToDevice, s));
  return NWK_OK;
}

int nwk_topk_state_update(nwk_topk_state* st, const float* dist, const int64_t* idx, void* stream) {
  if (!st || !dist || !idx) return fail(NWK_ERR_ARG, "null pointer argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)st->nq * st->k;
  NWK_CUDA(cudaMemcpyAsync(st->pack_d + n, dist, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  NWK_CUDA(cudaMemcpyAsync(st->pack_i + n, idx, n * sizeof(long long), cudaMemcpyDeviceToDevice, s));
  NWK_TRY(merge_topk(2, st->pack_d, st->pack_i, st->nq, st->k, st->k, st->tmp_d, st->tmp_i, st->ws, st->ws_bytes, s));
  return state_commit(st, s);
}

int nwk_topk_state_search(nwk_topk_state* st, const float* base, int64_t nb, const float* query, int32_t d,
                          int32_t metric, int64_t id_offset, void* workspace, size_t ws_bytes, uint32_t flags,
                          nwk_knn_stats* stats, void* stream) {
  if (!st) return fail(NWK_ERR_ARG, "null pointer argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)st->nq * st->k;
  // the slab's list goes straight into the incoming slot; the current k-th distances bound the search
  NWK_TRY(knn_impl(base, nb, query, st->nq, d, st->k, metric, id_offset, st->pack_d + n, (int64_t*)(st->pack_i + n),
                   workspace, ws_bytes, flags, stats, stream, st->pack_d + (st->k - 1), st->k));
  NWK_TRY(merge_topk(2, st->pack_d, st->pack_i, st->nq, st->k, st->k, st->tmp_d, st->tmp_i, st->ws, st->ws_bytes, s));
  return state_commit(st, s);
}

int nwk_topk_state_finalize(nwk_topk_state* st, float* out_dist, int64_t* out_idx, void* stream) {
  if (!st || !out_dist || !out_idx) return fail(NWK_ERR_ARG, "null pointer argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)st->nq * st->k;
  // (cudaMemcpyDefault: the destination may be device memory or pinned / pageable host memory)
  NWK_CUDA(cudaMemcpyAsync(out_dist, st->pack_d, n * sizeof(float), cudaMemcpyDefault, s));
  NWK_CUDA(cudaMemcpyAsync(out_idx, st->pack_i, n * sizeof(long long), cudaMemcpyDefault, s));
  return NWK_OK;
}

int nwk_topk_state_destroy(nwk_topk_state* st) {
  if (!st) return NWK_OK;
  cudaFree(st->pack_d);
  cudaFree(st->pack_i);
  cudaFree(st->tmp_d);
  cudaFree(st->tmp_i);
  cudaFree(st->ws);
  delete st;
  return NWK_OK;
}

// ---- host-buffer path -----------------------------------------------------------------

// Slab partition of nwk_knn_host_f32: slabs 0 .. nslab-2 hold slab_rows rows, the last one the
// remainder.  Every slab must hold at least k rows (the reference asserts the same of its last
// batch, cu_knn.py:177): a tail shorter than k is folded into the slab before it, so the last
// slab -- and the device ring -- holds fewer than slab_rows + k rows.
int nwk_host_slab_plan(int64_t nb, int64_t nq, int32_t d, int32_t k, int64_t slab_rows_in, int64_t* slab_rows_out,
                       int64_t* nslab_out, int64_t* last_rows_out) {
  if (nb <= 0 || nq <= 0 || d <= 0 || k <= 0 || k > nb || !slab_rows_out || !nslab_out || !last_rows_out)
    return fail(NWK_ERR_ARG, "bad slab plan arguments");
  long long slab_rows = slab_rows_in;
  if (slab_rows <= 0) {
    // Default: 8 slabs so the first copy is a small part of the run, but slabs of at least 64K rows so
    // the per-slab fixed cost stays negligible.  That fixed cost is per QUERY (a sample pass and a
    // finish stage per slab: 10-20 ms per slab at 100 000 queries, against 10 ms of exposed first
    // copy per GB), so large query sets take fewer, larger slabs; a slab is at most 6 GB.
    const long long want = nq > 50000 ? 3 : (nq > 20000 ? 4 : 8);
    slab_rows = (nb + want - 1) / want;
    if (slab_rows < 65536) slab_rows = 65536;
    const long long max_rows = (long long)((6ull << 30) / ((size_t)d * 4));
    if (slab_rows > max_rows) slab_rows = max_rows;
  }
  if (slab_rows > nb) slab_rows = nb;
  if (slab_rows < k) slab_rows = k;
  long long nslab = (nb + slab_rows - 1) / slab_rows;
  long long last = nb - (nslab - 1) * slab_rows;
  if (nslab > 1 && last < k) {
    nslab -= 1;
    last += slab_rows;
  }
  *slab_rows_out = slab_rows;
  *nslab_out = nslab;
  *last_rows_out = last;
  return NWK_OK;
}

int nwk_knn_host_f32(const float* base_host, int64_t nb, const float* query_host, int64_t nq, int32_t d,
                     int32_t k, int32_t metric, int64_t id_offset, int64_t slab_rows_arg, float* out_dist_host,
                     int64_t* out_idx_host, uint32_t flags, nwk_knn_stats* stats) {
  if (!base_host || !query_host || !out_dist_host || !out_idx_host) return fail(NWK_ERR_ARG, "null pointer argument");
  if (nb <= 0 || nq <= 0 || d <= 0 || k <= 0 || k > NWK_MAX_K || k > nb) return fail(NWK_ERR_ARG, "bad shape");
  if (metric < 0 || metric > 2) return fail(NWK_ERR_ARG, "unknown metric %d", metric);
  NWK_TRY(device_check());
  int dev = 0;
  cudaGetDevice(&dev);
  HostCtx& c = g_host[dev];
  std::lock_guard<std::mutex> lock(c.mu);
  if (!c.init) {
    NWK_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    NWK_CUDA(cudaStreamCreateWithFlags(&c.comp_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      NWK_CUDA(cudaEventCreateWithFlags(&c.copied[i], cudaEventDisableTiming));
      NWK_CUDA(cudaEventCreateWithFlags(&c.freed[i], cudaEventDisableTiming));
    }
    c.init = true;
  }
  int64_t slab_rows = 0, nslab = 0, last_rows = 0;
  NWK_TRY(nwk_host_slab_plan(nb, nq, d, k, slab_rows_arg, &slab_rows, &nslab, &last_rows));
  auto rows_of = [&](long long i) -> long long { return i + 1 < nslab ? slab_rows : last_rows; };
  const long long ring_rows = std::max<long long>(slab_rows, last_rows);
  const size_t n_out = (size_t)nq * k;
  size_t ws_need = nwk_knn_workspace_bytes(ring_rows, nq, d, k, metric, flags);
  if (nslab > 1) ws_need = std::max(ws_need, nwk_knn_workspace_bytes(slab_rows, nq, d, k, metric, flags));
  // slot of slab i: i & 1; the (longer) last slab needs ring_rows only in its own slot
  for (int b = 0; b < 2; ++b) {
    const bool holds_last = ((nslab - 1) & 1) == b;
    if (b == 1 && nslab == 1) break;
    NWK_TRY(grow(&c.slab[b], &c.slab_bytes[b], (size_t)(holds_last ? ring_rows : slab_rows) * d * 4));
  }
  NWK_TRY(grow(&c.query, &c.query_bytes, (size_t)nq * d * 4));
  NWK_TRY(grow(&c.ws, &c.ws_bytes, ws_need));
  nwk_topk_state* st = nullptr;
  if (nslab > 1) {
    if (c.state && (c.state->nq != nq || c.state->k != k)) {
      nwk_topk_state_destroy(c.state);
      c.state = nullptr;
    }
    if (!c.state) NWK_TRY(nwk_topk_state_create(nq, k, &c.state));
    st = c.state;
    NWK_TRY(nwk_topk_state_reset(st, c.comp_stream));
  } else {
    NWK_TRY(grow(&c.part_d, &c.part_d_bytes, n_out * sizeof(float)));
    NWK_TRY(grow(&c.part_i, &c.part_i_bytes, n_out * sizeof(long long)));
  }

  nwk_knn_stats total{};
  int rc = NWK_OK;
  auto enqueue_copy = [&](long long i) -> int {
    const int b = (int)(i & 1);
    NWK_CUDA(cudaStreamWaitEvent(c.copy_stream, c.freed[b], 0));
    NWK_CUDA(cudaMemcpyAsync(c.slab[b], base_host + (size_t)i * slab_rows * d, (size_t)rows_of(i) * d * 4,
                             cudaMemcpyHostToDevice, c.copy_stream));
    NWK_CUDA(cudaEventRecord(c.copied[b], c.copy_stream));
    return NWK_OK;
  };
  do {
    if ((rc = [&]() -> int {
          NWK_CUDA(cudaMemcpyAsync(c.query, query_host, (size_t)nq * d * 4, cudaMemcpyHostToDevice, c.comp_stream));
          // both ring slots start free
          NWK_CUDA(cudaEventRecord(c.freed[0], c.comp_stream));
          NWK_CUDA(cudaEventRecord(c.freed[1], c.comp_stream));
          return NWK_OK;
        }()) != NWK_OK) break;
    if ((rc = enqueue_copy(0)) != NWK_OK) break;
    for (long long i = 0; i < nslab && rc == NWK_OK; ++i) {
      const int b = (int)(i & 1);
      if (i + 1 < nslab && (rc = enqueue_copy(i + 1)) != NWK_OK) break;   // overlaps with this slab's search
      rc = [&]() -> int {
        NWK_CUDA(cudaStreamWaitEvent(c.comp_stream, c.copied[b], 0));
        nwk_knn_stats s1{};
        if (st)
          NWK_TRY(nwk_topk_state_search(st, c.slab[b], rows_of(i), c.query, d, metric, id_offset + i * slab_rows, c.ws,
                                        c.ws_bytes, flags, &s1, c.comp_stream));
        else
          NWK_TRY(nwk_knn_f32(c.slab[b], rows_of(i), c.query, nq, d, k, metric, id_offset + i * slab_rows, c.part_d,
                              (int64_t*)c.part_i, c.ws, c.ws_bytes, flags, &s1, c.comp_stream));
        NWK_CUDA(cudaEventRecord(c.freed[b], c.comp_stream));
        total.path = s1.path;
        total.n_fallback_queries += s1.n_fallback_queries;
        total.n_retry_queries += s1.n_retry_queries;
        total.total_candidates += s1.total_candidates;
        total.total_rerank += s1.total_rerank;
        total.gpu_launches += s1.gpu_launches;
        if (s1.max_candidates > total.max_candidates) total.max_candidates = s1.max_candidates;
        if (s1.max_rerank > total.max_rerank) total.max_rerank = s1.max_rerank;
        if (s1.max_err_ratio > total.max_err_ratio) total.max_err_ratio = s1.max_err_ratio;
        return NWK_OK;
      }();
    }
    if (rc != NWK_OK) break;
    rc = [&]() -> int {
      const float* fd = st ? st->pack_d : c.part_d;
      const long long* fi = st ? st->pack_i : c.part_i;
      NWK_CUDA(cudaMemcpyAsync(out_dist_host, fd, n_out * sizeof(float), cudaMemcpyDeviceToHost, c.comp_stream));
      NWK_CUDA(cudaMemcpyAsync(out_idx_host, fi, n_out * sizeof(long long), cudaMemcpyDeviceToHost, c.comp_stream));
      NWK_CUDA(cudaStreamSynchronize(c.comp_stream));
      return NWK_OK;
    }();
  } while (0);
  if (rc != NWK_OK) {
    cudaStreamSynchronize(c.copy_stream);
    cudaStreamSynchronize(c.comp_stream);
  }
  if (stats && rc == NWK_OK) *stats = total;
  return rc;
}

// Pinned host memory for the vectors the host path streams from.  write_combined: the pages are mapped
// write-combined on the CPU (slow for the CPU to read, but host -> device DMA reads of them need no cache
// snooping -- worth measuring where several GPUs pull from one host, see scripts/h2d_probe.py).
int nwk_host_alloc(size_t bytes, int32_t write_combined, void** out) {
  if (!out || bytes == 0) return fail(NWK_ERR_ARG, "bad host allocation request");
  NWK_TRY(device_check());
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0));
  if (e != cudaSuccess) return fail(NWK_ERR_CUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  *out = p;
  return NWK_OK;
}

int nwk_host_free(void* p) {
  if (!p) return NWK_OK;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) return fail(NWK_ERR_CUDA, "cudaFreeHost failed: %s", cudaGetErrorString(e));
  return NWK_OK;
}

// Frees what the host path has cached on the current device (staging ring, workspace, running state).
int nwk_host_release(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= NWK_MAX_DEVICES) return NWK_OK;
  HostCtx& c = g_host[dev];
  std::lock_guard<std::mutex> lock(c.mu);
  for (int b = 0; b < 2; ++b) { if (c.slab[b]) cudaFree(c.slab[b]); c.slab[b] = nullptr; c.slab_bytes[b] = 0; }
  if (c.query) cudaFree(c.query);
  if (c.ws) cudaFree(c.ws);
  if (c.part_d) cudaFree(c.part_d);
  if (c.part_i) cudaFree(c.part_i);
  c.query = nullptr; c.ws = nullptr; c.part_d = nullptr; c.part_i = nullptr;
  c.query_bytes = c.ws_bytes = c.part_d_bytes = c.part_i_bytes = 0;
  if (c.state) nwk_topk_state_destroy(c.state);
  c.state = nullptr;
  return NWK_OK;
}

size_t nwk_maxsim_workspace_bytes(int64_t nq, int32_t tq, int64_t nd, int32_t td, int32_t d, int32_t k) {
  if (nq <= 0 || nd <= 0 || tq <= 0 || td <= 0 || d <= 0 || k <= 0) return 256;
  return maxsim_workspace_bytes(nq, tq, nd, td, d, k);
}

int nwk_maxsim_f32(const float* query_tokens, int64_t nq, int32_t tq, const float* doc_tokens, int64_t nd, int32_t td,
                   const int32_t* doc_len, int32_t d, int32_t k, int64_t id_offset, float* out_score, int64_t* out_idx,
                   void* workspace, size_t ws_bytes, nwk_knn_stats* stats, void* stream) {
  if (!query_tokens || !doc_tokens || !out_score || !out_idx || !workspace) return fail(NWK_ERR_ARG, "null pointer argument");
  NWK_TRY(device_check());
  if (((reinterpret_cast<uintptr_t>(query_tokens) | reinterpret_cast<uintptr_t>(doc_tokens)) & 15u) != 0)
    return fail(NWK_ERR_ARG, "maxsim: token arrays must be 16-byte aligned");
  return maxsim_topk(query_tokens, nq, tq, doc_tokens, nd, td, doc_len, d, k, id_offset, out_score, (long long*)out_idx,
                     workspace, ws_bytes, stats, static_cast<cudaStream_t>(stream));
}

int nwk_synth_fill(float* out, int64_t rows, int32_t d, uint64_t seed, int64_t row0, int32_t normalize,
                   void* stream) {
  if (!out || rows < 0 || d <= 0) return fail(NWK_ERR_ARG, "bad synth arguments");
  NWK_TRY(device_check());
  return synth_fill(out, rows, d, seed, row0, normalize, static_cast<cudaStream_t>(stream));
}

int nwk_rows_from_columns_f32(const float* src, int64_t ld, int32_t d, int64_t rows, float* dst, void* stream) {
  if (!src || !dst || d <= 0 || rows < 0 || ld < rows) return fail(NWK_ERR_ARG, "bad layout arguments");
  NWK_TRY(device_check());
  return rows_from_columns(src, ld, d, rows, dst, static_cast<cudaStream_t>(stream));
}

int nwk_validate_pairs(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d, const int64_t* idx,
                       const float* dist, int32_t k, double rtol, double atol, int64_t* mismatches_host,
                       void* stream) {
  if (!base || !query || !idx || !dist || !mismatches_host) return fail(NWK_ERR_ARG, "null pointer argument");
  if (nb <= 0 || nq < 0 || d <= 0 || k <= 0) return fail(NWK_ERR_ARG, "bad shape");
  NWK_TRY(device_check());
  long long out = 0;
  NWK_TRY(validate_pairs(base, nb, query, nq, d, (const long long*)idx, dist, k, rtol, atol, &out,
                         static_cast<cudaStream_t>(stream)));
  *mismatches_host = out;
  return NWK_OK;
}

int nwk_debug_tc_tile(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d, int64_t b0,
                      int64_t q0, float* out, int32_t use_f16, void* stream) {
  if (!base || !query || !out) return fail(NWK_ERR_ARG, "null pointer argument");
  NWK_TRY(device_check());
  return tc_debug_tile(base, nb, query, nq, d, b0, q0, out, use_f16, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---- single-process multi-GPU ---------------------------------------------------------
// nwk_mg_*: one host process drives n devices (a host thread per device for the duration of a call).
// The base is row-sharded: shard i lives on devices[i]; the queries come from host memory and are
// replicated; every device computes its exact local top-k with global ids; device 0 then merges the n
// lists reading them straight out of the peers' memory (cudaDeviceEnablePeerAccess: the merge kernel's
// loads cross NVLink -- gather and merge are one kernel, no staging copy).  The multi-process twin,
// for one process per GPU, is nwk_xchg_* (nwk_xchg.cu).
#include <thread>
#include <vector>

struct nwk_mg {
  int n = 0;
  int dev[16] = {};
  cudaStream_t stream[16] = {};
  float* query[16] = {};
  size_t query_bytes[16] = {};
  char* ws[16] = {};
  size_t ws_bytes[16] = {};
  float* out_d[16] = {};
  size_t out_d_bytes[16] = {};
  long long* out_i[16] = {};
  size_t out_i_bytes[16] = {};
  float* fin_d = nullptr;          // device 0
  size_t fin_d_bytes = 0;
  long long* fin_i = nullptr;
  size_t fin_i_bytes = 0;
  char* merge_ws = nullptr;
  size_t merge_ws_bytes = 0;
};

extern "C" {

int nwk_mg_create(const int32_t* devices, int32_t n, nwk_mg** out) {
  if (!devices || !out || n < 1 || n > 16) return fail(NWK_ERR_ARG, "nwk_mg_create: 1..16 devices");
  int prev = 0;
  cudaGetDevice(&prev);
  nwk_mg* mg = new nwk_mg();
  mg->n = n;
  int rc = NWK_OK;
  for (int i = 0; i < n && rc == NWK_OK; ++i) {
    mg->dev[i] = devices[i];
    if (cudaSetDevice(devices[i]) != cudaSuccess) { rc = fail(NWK_ERR_DEVICE, "cannot select device %d", devices[i]); break; }
    rc = device_check();
    if (rc == NWK_OK && cudaStreamCreateWithFlags(&mg->stream[i], cudaStreamNonBlocking) != cudaSuccess)
      rc = fail(NWK_ERR_CUDA, "stream creation failed on device %d", devices[i]);
  }
  // device 0 reads the peers' result lists
  if (rc == NWK_OK) {
    cudaSetDevice(devices[0]);
    for (int i = 1; i < n; ++i) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[0], devices[i]);
      if (!can) { rc = fail(NWK_ERR_DEVICE, "device %d cannot access device %d's memory", devices[0], devices[i]); break; }
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = fail(NWK_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); break; }
      cudaGetLastError();
    }
  }
  cudaSetDevice(prev);
  if (rc != NWK_OK) { nwk_mg_destroy(mg); return rc; }
  *out = mg;
  return NWK_OK;
}

int nwk_mg_knn_f32(nwk_mg* mg, const float* const* base_shards, const int64_t* rows, const float* query_host, int64_t nq,
                   int32_t d, int32_t k, int32_t metric, int64_t id_offset, float* out_dist_host, int64_t* out_idx_host,
                   uint32_t flags, nwk_knn_stats* stats) {
  if (!mg || !base_shards || !rows || !query_host || !out_dist_host || !out_idx_host) return fail(NWK_ERR_ARG, "null pointer argument");
  if (nq <= 0 || d <= 0 || k <= 0 || k > NWK_MAX_K || (long long)mg->n * k > 2048) return fail(NWK_ERR_ARG, "bad shape");
  for (int i = 0; i < mg->n; ++i)
    if (!base_shards[i] || rows[i] < k) return fail(NWK_ERR_ARG, "shard %d: every shard must hold at least k rows", i);
  int prev = 0;
  cudaGetDevice(&prev);
  const int n = mg->n;
  std::vector<int> rcs(n, NWK_OK);
  std::vector<std::string> msgs(n);
  std::vector<nwk_knn_stats> sts(n);
  std::vector<long long> offs(n, id_offset);
  for (int i = 1; i < n; ++i) offs[i] = offs[i - 1] + rows[i - 1];
  const size_t n_out = (size_t)nq * k;
  auto worker = [&](int i) {
    auto run = [&]() -> int {
      NWK_CUDA(cudaSetDevice(mg->dev[i]));
      NWK_TRY(device_check());
      const size_t need = nwk_knn_workspace_bytes(rows[i], nq, d, k, metric, flags);
      NWK_TRY(grow(&mg->query[i], &mg->query_bytes[i], (size_t)nq * d * 4));
      NWK_TRY(grow(&mg->ws[i], &mg->ws_bytes[i], need));
      NWK_TRY(grow(&mg->out_d[i], &mg->out_d_bytes[i], n_out * sizeof(float)));
      NWK_TRY(grow(&mg->out_i[i], &mg->out_i_bytes[i], n_out * sizeof(long long)));
      NWK_CUDA(cudaMemcpyAsync(mg->query[i], query_host, (size_t)nq * d * 4, cudaMemcpyHostToDevice, mg->stream[i]));
      NWK_TRY(nwk_knn_f32(base_shards[i], rows[i], mg->query[i], nq, d, k, metric, offs[i], mg->out_d[i],
                          (int64_t*)mg->out_i[i], mg->ws[i], mg->ws_bytes[i], flags, &sts[i], mg->stream[i]));
      NWK_CUDA(cudaStreamSynchronize(mg->stream[i]));
      return NWK_OK;
    };
    rcs[i] = run();
    if (rcs[i] != NWK_OK) msgs[i] = last_error_ref();
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i) th.emplace_back(worker, i);
  worker(0);
  for (auto& t : th) t.join();
  cudaSetDevice(mg->dev[0]);
  for (int i = 0; i < n; ++i)
    if (rcs[i] != NWK_OK) { cudaSetDevice(prev); return fail(rcs[i], "device %d: %s", mg->dev[i], msgs[i].c_str()); }
  int rc = [&]() -> int {
    const float* dp[16];
    const long long* ip[16];
    for (int i = 0; i < n; ++i) { dp[i] = mg->out_d[i]; ip[i] = mg->out_i[i]; }
    const float* fd = mg->out_d[0];
    const long long* fi = mg->out_i[0];
    if (n > 1) {
      NWK_TRY(grow(&mg->fin_d, &mg->fin_d_bytes, n_out * sizeof(float)));
      NWK_TRY(grow(&mg->fin_i, &mg->fin_i_bytes, n_out * sizeof(long long)));
      NWK_TRY(grow(&mg->merge_ws, &mg->merge_ws_bytes, nwk_merge_workspace_bytes(n, nq, k, k)));
      // gather + merge in one kernel: the lists of devices 1.. are read over NVLink
      NWK_TRY(merge_topk_ptrs(n, dp, ip, nq, k, k, mg->fin_d, mg->fin_i, mg->merge_ws, mg->merge_ws_bytes, mg->stream[0]));
      fd = mg->fin_d;
      fi = mg->fin_i;
    }
    NWK_CUDA(cudaMemcpyAsync(out_dist_host, fd, n_out * sizeof(float), cudaMemcpyDeviceToHost, mg->stream[0]));
    NWK_CUDA(cudaMemcpyAsync(out_idx_host, fi, n_out * sizeof(long long), cudaMemcpyDeviceToHost, mg->stream[0]));
    NWK_CUDA(cudaStreamSynchronize(mg->stream[0]));
    return NWK_OK;
  }();
  cudaSetDevice(prev);
  if (rc == NWK_OK && stats) {
    nwk_knn_stats t = sts[0];
    for (int i = 1; i < n; ++i) {
      t.n_fallback_queries += sts[i].n_fallback_queries;
      t.n_retry_queries += sts[i].n_retry_queries;
      t.total_candidates += sts[i].total_candidates;
      t.total_rerank += sts[i].total_rerank;
      t.gpu_launches += sts[i].gpu_launches;
      t.max_candidates = std::max(t.max_candidates, sts[i].max_candidates);
      t.max_rerank = std::max(t.max_rerank, sts[i].max_rerank);
      t.max_err_ratio = std::max(t.max_err_ratio, sts[i].max_err_ratio);
    }
    *stats = t;
  }
  return rc;
}

int nwk_mg_destroy(nwk_mg* mg) {
  if (!mg) return NWK_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int i = 0; i < mg->n; ++i) {
    if (cudaSetDevice(mg->dev[i]) != cudaSuccess) continue;
    if (mg->stream[i]) { cudaStreamSynchronize(mg->stream[i]); cudaStreamDestroy(mg->stream[i]); }
    cudaFree(mg->query[i]); cudaFree(mg->ws[i]); cudaFree(mg->out_d[i]); cudaFree(mg->out_i[i]);
    if (i == 0) { cudaFree(mg->fin_d); cudaFree(mg->fin_i); cudaFree(mg->merge_ws); }
  }
  cudaSetDevice(prev);
  delete mg;
  return NWK_OK;
}

}  // extern "C"
