// nwk_scan.cu -- fp64 exact scan, segmented top-k and the k-way merge.
//
// The exact scan is the engine's ground floor: every (query, base row) distance is
// evaluated in fp64 straight from the fp32 inputs and the k smallest are selected by
// (distance, index).  It serves three purposes:
//   * the fallback for queries whose tensor-core candidate set cannot be certified,
//   * the whole path for shapes the tcgen05 kernel does not take (d % 4 != 0, tiny inputs),
//   * the on-device verifier the large parity tests sample against.
// It replaces the same reference call sites as nwk_knn_f32 (cu_knn.py:243-257).
//
// The merge kernel replaces the argsort fold of merge.py:73-104.
#include "nwk_common.cuh"

namespace nwk {

constexpr int SCAN_QG = 8;        // queries per CTA of the distance kernel
constexpr int SCAN_WARPS = 8;
constexpr int SEG = 4096;         // elements sorted per CTA of the selection kernels
constexpr int SEG_THREADS = 512;

// ---------------------------------------------------------------------------------------
// distance kernel: dist64[f][row] for a group of SCAN_QG selected queries
//   HBM-bound only when few queries are selected; otherwise fp64-pipe bound.
//   bytes per launch: nb*d*4 (base read once per query group) + F*nb*8 written.
// ---------------------------------------------------------------------------------------
template <int METRIC>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_dist_kernel(const float* __restrict__ base, long long nb, const float* __restrict__ query,
                 int d, const int* __restrict__ qsel, int f0, int nsel,
                 double* __restrict__ dist64) {
  extern __shared__ float s_q[];                       // [SCAN_QG][d]
  __shared__ double s_qn2[SCAN_QG];
  const int g0 = f0 + blockIdx.y * SCAN_QG;            // first selected slot of this group
  const int ng = max(0, min(SCAN_QG, f0 + nsel - g0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < SCAN_QG * d; i += blockDim.x) {
    int g = i / d, c = i - g * d;
    float v = 0.f;
    if (g < ng) {
      long long qrow = qsel ? qsel[g0 + g] : (long long)(g0 + g);
      v = query[qrow * d + c];
    }
    s_q[i] = v;
  }
  __syncthreads();
  if (METRIC == NWK_METRIC_COSINE && warp < SCAN_QG) {
    double a = 0.0;
    for (int c = lane; c < d; c += 32) { double v = s_q[warp * d + c]; a = fma(v, v, a); }
    a = warp_sum(a);
    if (lane == 0) s_qn2[warp] = a;
  }
  __syncthreads();

  const bool vec4 = (d & 3) == 0;
  for (long long row = (long long)blockIdx.x * SCAN_WARPS + warp; row < nb;
       row += (long long)gridDim.x * SCAN_WARPS) {
    double acc[SCAN_QG];
#pragma unroll
    for (int g = 0; g < SCAN_QG; ++g) acc[g] = 0.0;
    double bn2 = 0.0;
    const float* brow = base + row * d;
    if (vec4) {
      for (int c = lane * 4; c < d; c += 128) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c));
        double b0 = b4.x, b1 = b4.y, b2 = b4.z, b3 = b4.w;
        if (METRIC == NWK_METRIC_COSINE) {
          bn2 = fma(b0, b0, bn2); bn2 = fma(b1, b1, bn2); bn2 = fma(b2, b2, bn2); bn2 = fma(b3, b3, bn2);
        }
#pragma unroll
        for (int g = 0; g < SCAN_QG; ++g) {
          float4 q4 = *reinterpret_cast<const float4*>(s_q + g * d + c);
          if (METRIC == NWK_METRIC_SQEUCLIDEAN) {
            double e0 = (double)q4.x - b0, e1 = (double)q4.y - b1;
            double e2 = (double)q4.z - b2, e3 = (double)q4.w - b3;
            acc[g] = fma(e0, e0, acc[g]); acc[g] = fma(e1, e1, acc[g]);
            acc[g] = fma(e2, e2, acc[g]); acc[g] = fma(e3, e3, acc[g]);
          } else {
            // operation order shared with warp_exact_distance (nwk_tc.cu): bit-identical fp64
            acc[g] = fma((double)q4.x, b0, acc[g]); acc[g] = fma((double)q4.y, b1, acc[g]);
            acc[g] = fma((double)q4.z, b2, acc[g]); acc[g] = fma((double)q4.w, b3, acc[g]);
          }
        }
      }
    } else {
      for (int c = lane; c < d; c += 32) {
        double b0 = __ldg(brow + c);
        if (METRIC == NWK_METRIC_COSINE) bn2 = fma(b0, b0, bn2);
#pragma unroll
        for (int g = 0; g < SCAN_QG; ++g) {
          double qv = s_q[g * d + c];
          if (METRIC == NWK_METRIC_SQEUCLIDEAN) { double e = qv - b0; acc[g] = fma(e, e, acc[g]); }
          else acc[g] = fma(qv, b0, acc[g]);
        }
      }
    }
    if (METRIC == NWK_METRIC_COSINE) bn2 = warp_sum(bn2);
#pragma unroll
    for (int g = 0; g < SCAN_QG; ++g) acc[g] = warp_sum(acc[g]);
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < SCAN_QG; ++g) {
        if (g < ng) {
          double v = acc[g];
          if (METRIC != NWK_METRIC_SQEUCLIDEAN) v = finish_metric(METRIC, v, s_qn2[g], bn2);
          dist64[(long long)(g0 - f0 + g) * nb + row] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// segmented top-k: each CTA sorts one segment of up to SEG (key, id) pairs of one query
// and emits its k smallest.  Sources:
//   SRC_IOTA   keys double [F][n], id = position           (first round of the exact scan)
//   SRC_PAIRS  keys double + ids int64 [F][n]              (later rounds)
//   SRC_MERGE  keys float [P][nq][k_in] + ids int64        (first round of the merge)
// When `final_out` is set the k results go to (out_dist fp32, out_idx int64 + id_offset) at
// row `qsel ? qsel[f] : f`, otherwise to (out_key, out_id) [F][nseg*k].
// ---------------------------------------------------------------------------------------
enum { SRC_IOTA = 0, SRC_PAIRS = 1, SRC_MERGE = 2 };

// Where the P partial lists of a merge live: list p is [nq][k_in] row-major at dist_tab[p] / idx_tab[p]
// (P <= MERGE_TAB pointers: lists in the memory of P peer GPUs, or in separately allocated buffers), or
// at dist + p * dist_stride / idx + p * idx_stride (any P: one contiguous [P, nq, k_in] array, or the
// chunks of a packed all-gather buffer).
constexpr int MERGE_TAB = 16;
struct MergeSrc {
  const float* dist;
  const long long* idx;
  long long dist_stride, idx_stride;        // elements between consecutive lists
  int use_tab;
  const float* dist_tab[MERGE_TAB];
  const long long* idx_tab[MERGE_TAB];
  __device__ __forceinline__ const float* d(int p) const { return use_tab ? dist_tab[p] : dist + p * dist_stride; }
  __device__ __forceinline__ const long long* i(int p) const { return use_tab ? idx_tab[p] : idx + p * idx_stride; }
};

// Padding of a partial list is marked by a negative id (-1), whatever its distance: it orders after
// every real entry -- including real entries whose distance is +inf or nan (nan sorts as +inf) -- and
// comes out as (+inf, -1) again.
constexpr long long PAD_ID = 0x7fffffffffffffffLL;

struct SegArgs {
  const double* in_key;
  const long long* in_id;
  MergeSrc src;             // SRC_MERGE
  long long q0;             // SRC_MERGE: first query of this launch
  long long n;              // elements per query
  long long in_stride;      // stride between queries (elements) for SRC_IOTA / SRC_PAIRS
  int k_in;                 // SRC_MERGE
  int k;
  double* out_key;
  long long* out_id;
  long long out_stride;     // nseg * k
  float* out_dist;
  long long* out_idx;
  long long id_offset;
  const int* qsel;
  int f0;                   // slot offset into qsel for this launch
  int final_out;
};

template <int SRC>
__global__ void __launch_bounds__(SEG_THREADS)
seg_topk_kernel(SegArgs a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  double* keys = reinterpret_cast<double*>(s_raw);
  long long* ids = reinterpret_cast<long long*>(s_raw + sizeof(double) * SEG);
  const long long f = blockIdx.y;
  const long long seg = blockIdx.x;
  const long long e0 = seg * SEG;
  const int valid = (int)min((long long)SEG, a.n - e0);
  const int n2 = max(next_pow2(valid), 2);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  for (int t = threadIdx.x; t < n2; t += blockDim.x) {
    double kv = INF;
    long long iv = 0x7fffffffffffffffLL;
    if (t < valid) {
      long long e = e0 + t;
      if (SRC == SRC_IOTA) {
        kv = a.in_key[f * a.in_stride + e];
        iv = e;
      } else if (SRC == SRC_PAIRS) {
        kv = a.in_key[f * a.in_stride + e];
        iv = a.in_id[f * a.in_stride + e];
      } else {
        const int p = (int)(e / a.k_in);
        const long long off = (a.q0 + f) * a.k_in + (e - (long long)p * a.k_in);
        kv = (double)a.src.d(p)[off];
        iv = a.src.i(p)[off];
        if (iv < 0) { kv = INF; iv = PAD_ID; }
      }
      if (kv != kv) kv = INF;   // NaN sorts last
    }
    keys[t] = kv;
    ids[t] = iv;
  }
  block_bitonic_sort(keys, ids, n2);

  for (int t = threadIdx.x; t < a.k; t += blockDim.x) {
    double kv = t < n2 ? keys[t] : INF;
    long long iv = t < n2 ? ids[t] : 0x7fffffffffffffffLL;
    if (a.final_out) {
      long long row = a.qsel ? (long long)a.qsel[a.f0 + f] : f;
      a.out_dist[row * a.k + t] = (float)kv;
      a.out_idx[row * a.k + t] = (iv == 0x7fffffffffffffffLL) ? -1 : iv + a.id_offset;
    } else {
      a.out_key[f * a.out_stride + seg * a.k + t] = kv;
      a.out_id[f * a.out_stride + seg * a.k + t] = iv;
    }
  }
}

static int seg_smem_optin() {
  DevState* ds = dev_state();
  if (!ds->seg_optin) {
    const int bytes = SEG * 16;
    cudaError_t e;
    e = cudaFuncSetAttribute(seg_topk_kernel<SRC_IOTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return fail(NWK_ERR_CUDA, "smem opt-in failed: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(seg_topk_kernel<SRC_PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return fail(NWK_ERR_CUDA, "smem opt-in failed: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(seg_topk_kernel<SRC_MERGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return fail(NWK_ERR_CUDA, "smem opt-in failed: %s", cudaGetErrorString(e));
    ds->seg_optin = 1;
  }
  return NWK_OK;
}

// Reduce [F][n] candidates (described by `a`, first-round source SRC) to the final k per query.
// ping / pong: scratch of F * ceil(n/SEG) * k pairs each.
template <int SRC>
static int run_topk_rounds(SegArgs a, int F, double* ping_k, long long* ping_i, double* pong_k,
                           long long* pong_i, cudaStream_t s) {
  NWK_TRY(seg_smem_optin());
  const int smem = SEG * 16;
  long long n = a.n;
  bool first = true;
  float* out_dist = a.out_dist;
  long long* out_idx = a.out_idx;
  while (true) {
    long long nseg = (n + SEG - 1) / SEG;
    a.n = n;
    a.final_out = (nseg == 1);
    a.out_key = ping_k;
    a.out_id = ping_i;
    a.out_stride = nseg * a.k;
    a.out_dist = out_dist;
    a.out_idx = out_idx;
    dim3 grid((unsigned)nseg, (unsigned)F);
    if (first) seg_topk_kernel<SRC><<<grid, SEG_THREADS, smem, s>>>(a);
    else seg_topk_kernel<SRC_PAIRS><<<grid, SEG_THREADS, smem, s>>>(a);
    NWK_LAUNCH_CHECK();
    if (nseg == 1) break;
    // next round reads what this round wrote
    a.in_key = ping_k;
    a.in_id = ping_i;
    a.in_stride = nseg * a.k;
    n = nseg * a.k;
    first = false;
    double* tk = ping_k; ping_k = pong_k; pong_k = tk;
    long long* ti = ping_i; ping_i = pong_i; pong_i = ti;
  }
  return NWK_OK;
}

// ---------------------------------------------------------------------------------------
// k smallest of F rows of n double keys (id = position + id_offset), for callers outside this
// file (the MaxSim path): out_key fp32 / out_idx int64 at row `qsel ? qsel[f] : f`.
// ---------------------------------------------------------------------------------------
size_t topk_from_keys_workspace_bytes(long long n, long long F, int k) {
  const long long nseg = (n + SEG - 1) / SEG;
  if (F > 32768) F = 32768;
  return nseg > 1 ? (size_t)F * nseg * k * 16 * 2 + 4096 : 256;
}

int topk_from_keys(const double* keys, long long n, long long stride, long long F, int k, const int* qsel,
                   long long id_offset, float* out_key, long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (F == 0) return NWK_OK;
  if (ws_bytes < topk_from_keys_workspace_bytes(n, F, k)) return fail(NWK_ERR_WORKSPACE, "top-k workspace too small");
  const long long nseg = (n + SEG - 1) / SEG;
  const long long SLICE = 32768;                       // grid.y limit
  const long long fc = F < SLICE ? F : SLICE;
  Arena ar(ws, ws_bytes);
  double* ping_k = nullptr; long long* ping_i = nullptr; double* pong_k = nullptr; long long* pong_i = nullptr;
  if (nseg > 1) {
    ping_k = ar.take<double>((size_t)fc * nseg * k);
    ping_i = ar.take<long long>((size_t)fc * nseg * k);
    pong_k = ar.take<double>((size_t)fc * nseg * k);
    pong_i = ar.take<long long>((size_t)fc * nseg * k);
  }
  for (long long f0 = 0; f0 < F; f0 += SLICE) {
    const int nf = (int)min(SLICE, F - f0);
    SegArgs a{};
    a.in_key = keys + f0 * stride;
    a.n = n;
    a.in_stride = stride;
    a.k = k;
    a.out_dist = out_key;
    a.out_idx = out_idx;
    a.id_offset = id_offset;
    a.qsel = qsel;
    a.f0 = (int)f0;
    if (!qsel) { a.out_dist = out_key + f0 * k; a.out_idx = out_idx + f0 * k; }
    NWK_TRY(run_topk_rounds<SRC_IOTA>(a, nf, ping_k, ping_i, pong_k, pong_i, s));
  }
  return NWK_OK;
}

// ---------------------------------------------------------------------------------------
// exact scan driver
// ---------------------------------------------------------------------------------------
static long long scan_chunk_queries(long long nb) {
  const long long budget = 512ll << 20;
  long long fc = budget / (nb * 8);
  fc = fc / SCAN_QG * SCAN_QG;
  if (fc < SCAN_QG) fc = SCAN_QG;
  if (fc > 4096) fc = 4096;
  return fc;
}

size_t exact_scan_workspace_bytes(long long nb, long long nsel, int k) {
  long long fc = scan_chunk_queries(nb);
  if (fc > nsel) fc = (nsel + SCAN_QG - 1) / SCAN_QG * SCAN_QG;
  if (fc < SCAN_QG) fc = SCAN_QG;
  long long nseg = (nb + SEG - 1) / SEG;
  size_t bytes = 0;
  bytes += align_up((size_t)fc * nb * 8, 256);
  bytes += 2 * (align_up((size_t)fc * nseg * k * 8, 256) * 2);
  return bytes + 4096;
}

// qsel: device list of query rows (nullptr = 0..nsel-1).  Results are written to row
// qsel[i] of out_dist / out_idx.
int exact_scan(const float* base, long long nb, const float* query, int d, int k, int metric,
               long long id_offset, const int* qsel, long long nsel, float* out_dist,
               long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (nsel == 0) return NWK_OK;
  if (ws_bytes < exact_scan_workspace_bytes(nb, nsel, k))
    return fail(NWK_ERR_WORKSPACE, "exact scan workspace too small");
  long long fc = scan_chunk_queries(nb);
  if (fc > nsel) fc = (nsel + SCAN_QG - 1) / SCAN_QG * SCAN_QG;
  long long nseg = (nb + SEG - 1) / SEG;
  Arena ar(ws, ws_bytes);
  double* dist64 = ar.take<double>((size_t)fc * nb);
  double* ping_k = ar.take<double>((size_t)fc * nseg * k);
  long long* ping_i = ar.take<long long>((size_t)fc * nseg * k);
  double* pong_k = ar.take<double>((size_t)fc * nseg * k);
  long long* pong_i = ar.take<long long>((size_t)fc * nseg * k);

  const size_t q_smem = (size_t)SCAN_QG * d * sizeof(float);
  if (q_smem > 200 * 1024) return fail(NWK_ERR_ARG, "d=%d too large for the exact scan", d);
  DevState* ds = dev_state();
  if (!ds->scan_optin) {
    NWK_CUDA(cudaFuncSetAttribute(scan_dist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    NWK_CUDA(cudaFuncSetAttribute(scan_dist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    NWK_CUDA(cudaFuncSetAttribute(scan_dist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ds->scan_optin = 1;
  }
  const int sms = sm_count();
  for (long long f0 = 0; f0 < nsel; f0 += fc) {
    const int nf = (int)min(fc, nsel - f0);
    const int groups = (nf + SCAN_QG - 1) / SCAN_QG;
    long long row_blocks = (nb + SCAN_WARPS - 1) / SCAN_WARPS;
    long long want = (long long)sms * 8 / groups;       // ~8 CTAs per SM in total
    if (want < 1) want = 1;
    if (row_blocks > want) row_blocks = want;
    dim3 grid((unsigned)row_blocks, (unsigned)groups);
    switch (metric) {
      case NWK_METRIC_SQEUCLIDEAN:
        scan_dist_kernel<NWK_METRIC_SQEUCLIDEAN><<<grid, SCAN_WARPS * 32, q_smem, s>>>(
            base, nb, query, d, qsel, (int)f0, nf, dist64);
        break;
      case NWK_METRIC_COSINE:
        scan_dist_kernel<NWK_METRIC_COSINE><<<grid, SCAN_WARPS * 32, q_smem, s>>>(
            base, nb, query, d, qsel, (int)f0, nf, dist64);
        break;
      default:
        scan_dist_kernel<NWK_METRIC_DOT><<<grid, SCAN_WARPS * 32, q_smem, s>>>(
            base, nb, query, d, qsel, (int)f0, nf, dist64);
        break;
    }
    NWK_LAUNCH_CHECK();
    SegArgs a{};
    a.in_key = dist64;
    a.n = nb;
    a.in_stride = nb;
    a.k = k;
    a.out_dist = out_dist;
    a.out_idx = out_idx;
    a.id_offset = id_offset;
    a.qsel = qsel;
    a.f0 = (int)f0;
    if (!qsel) {
      // identity selection: rows f0 .. f0+nf-1
      a.out_dist = out_dist + f0 * k;
      a.out_idx = out_idx + f0 * k;
    }
    NWK_TRY(run_topk_rounds<SRC_IOTA>(a, nf, ping_k, ping_i, pong_k, pong_i, s));
  }
  return NWK_OK;
}

// ---------------------------------------------------------------------------------------
// merge driver
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// rank merge: P lists of k_in records per query, P * k_in <= RANK_MERGE_MAX.  One warp per query
// with the lists in shared memory.  Lists that are non-decreasing in distance (what knn() and
// the running state produce) are merged by ranking: an element's slot in the output is the number
// of elements ordered before it, found by one binary search per list plus a scan over the (rare)
// equal distances; +inf entries are padding and stay out.  A query whose lists are not sorted is
// ranked by counting over all pairs instead.  Order: (distance, index, list, position) -- the
// first two are the documented order, the last two only separate exact duplicates.  NaN sorts
// as +inf.  (The segmented bitonic sort this replaces for small merges took 0.9 ms for 10 000
// queries x 8 lists x 100; the all-gather + merge of an 8-GPU step 1.37 ms.)
// ---------------------------------------------------------------------------------------
constexpr int RANK_MERGE_MAX = 2048;
constexpr int RANK_MERGE_THREADS = 128;

__global__ void __launch_bounds__(RANK_MERGE_THREADS)
merge_rank_kernel(const MergeSrc src, int P, long long nq, int k_in, int k_out, float* __restrict__ out_dist,
                  long long* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char rm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = P * k_in;
  long long* s_id = reinterpret_cast<long long*>(rm_smem) + (size_t)warp * n;
  float* s_d = reinterpret_cast<float*>(rm_smem + (size_t)(RANK_MERGE_THREADS / 32) * n * 8) + (size_t)warp * n;
  const long long q = (long long)blockIdx.x * (RANK_MERGE_THREADS / 32) + warp;
  if (q >= nq) return;
  const float INF = __int_as_float(0x7f800000);
  bool sorted = true;
  for (int e = lane; e < n; e += 32) {
    const int b = e / k_in, j = e - b * k_in;
    const size_t g = (size_t)q * k_in + j;
    float d = src.d(b)[g];
    long long id = src.i(b)[g];
    if (!(d == d)) d = INF;
    if (id < 0) { d = INF; id = PAD_ID; }           // padding: after every real entry
    s_d[e] = d;
    s_id[e] = id;
  }
  __syncwarp();
  for (int e = lane; e < n; e += 32)
    if ((e % k_in) != 0 && s_d[e - 1] > s_d[e]) sorted = false;
  sorted = __all_sync(0xffffffffu, sorted);
  // (distance, id, position) is a strict total order over the n entries, so the ranks are a
  // permutation of 0 .. n-1 and every output slot below min(k_out, n) has exactly one writer
  for (int e = lane; e < n; e += 32) {
    const float d = s_d[e];
    const long long id = s_id[e];
    int rank = 0;
    if (sorted) {
      for (int b = 0; b < P; ++b) {
        const float* l = s_d + b * k_in;
        int lo = 0, hi = k_in;                      // first position with l[pos] >= d
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
    if (rank < k_out) { out_dist[q * k_out + rank] = d; out_idx[q * k_out + rank] = (id == PAD_ID) ? -1 : id; }
  }
}

static int merge_rank(const MergeSrc& src, int P, long long nq, int k_in, int k_out, float* out_dist, long long* out_idx,
                      cudaStream_t s) {
  const int n = P * k_in;
  const int smem = (RANK_MERGE_THREADS / 32) * n * 12;
  DevState* ds = dev_state();
  if (smem > ds->merge_rank_optin) {
    NWK_CUDA(cudaFuncSetAttribute(merge_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ds->merge_rank_optin = smem;
  }
  const long long blocks = (nq + RANK_MERGE_THREADS / 32 - 1) / (RANK_MERGE_THREADS / 32);
  merge_rank_kernel<<<(unsigned)blocks, RANK_MERGE_THREADS, smem, s>>>(src, P, nq, k_in, k_out, out_dist, out_idx);
  NWK_LAUNCH_CHECK();
  return NWK_OK;
}

size_t merge_workspace_bytes(int P, long long nq, int k_in, int k_out) {
  long long n = (long long)P * k_in;
  long long nseg = (n + SEG - 1) / SEG;
  if (nseg <= 1) return 256;
  return 2 * (align_up((size_t)nq * nseg * k_out * 8, 256) * 2) + 4096;
}

int merge_topk_src(const MergeSrc& src, int P, long long nq, int k_in, int k_out, float* out_dist, long long* out_idx,
                   void* ws, size_t ws_bytes, cudaStream_t s) {
  if (nq == 0) return NWK_OK;
  if (ws_bytes < merge_workspace_bytes(P, nq, k_in, k_out))
    return fail(NWK_ERR_WORKSPACE, "merge workspace too small");
  if ((long long)P * k_in <= RANK_MERGE_MAX && nq < (1ll << 31) * (RANK_MERGE_THREADS / 32))
    return merge_rank(src, P, nq, k_in, k_out, out_dist, out_idx, s);
  // segmented path: every round keeps k_out of each 4096-record segment, so k_out <= SEG / 2 converges
  if (k_out > SEG / 2) return fail(NWK_ERR_ARG, "k_out=%d exceeds %d", k_out, SEG / 2);
  long long n = (long long)P * k_in;
  long long nseg = (n + SEG - 1) / SEG;
  Arena ar(ws, ws_bytes);
  double* ping_k = nullptr; long long* ping_i = nullptr;
  double* pong_k = nullptr; long long* pong_i = nullptr;
  if (nseg > 1) {
    ping_k = ar.take<double>((size_t)nq * nseg * k_out);
    ping_i = ar.take<long long>((size_t)nq * nseg * k_out);
    pong_k = ar.take<double>((size_t)nq * nseg * k_out);
    pong_i = ar.take<long long>((size_t)nq * nseg * k_out);
  }
  // grid.y is limited to 65535: walk the queries in slices
  const long long SLICE = 32768;
  for (long long q0 = 0; q0 < nq; q0 += SLICE) {
    long long nf = min(SLICE, nq - q0);
    SegArgs a{};
    a.src = src;
    a.q0 = q0;
    a.n = n;
    a.k_in = k_in;
    a.k = k_out;
    a.out_dist = out_dist + q0 * k_out;
    a.out_idx = out_idx + q0 * k_out;
    a.id_offset = 0;
    NWK_TRY(run_topk_rounds<SRC_MERGE>(a, (int)nf, ping_k, ping_i, pong_k, pong_i, s));
  }
  return NWK_OK;
}

int merge_topk_strided(int P, const float* dist, long long dist_stride, const long long* idx, long long idx_stride,
                       long long nq, int k_in, int k_out, float* out_dist, long long* out_idx, void* ws,
                       size_t ws_bytes, cudaStream_t s) {
  MergeSrc src{};
  src.dist = dist; src.idx = idx; src.dist_stride = dist_stride; src.idx_stride = idx_stride; src.use_tab = 0;
  return merge_topk_src(src, P, nq, k_in, k_out, out_dist, out_idx, ws, ws_bytes, s);
}

int merge_topk_ptrs(int P, const float* const* dist, const long long* const* idx, long long nq, int k_in, int k_out,
                    float* out_dist, long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (P > MERGE_TAB) return fail(NWK_ERR_ARG, "at most %d separately placed lists per merge", MERGE_TAB);
  MergeSrc src{};
  src.use_tab = 1;
  for (int p = 0; p < P; ++p) { src.dist_tab[p] = dist[p]; src.idx_tab[p] = idx[p]; }
  return merge_topk_src(src, P, nq, k_in, k_out, out_dist, out_idx, ws, ws_bytes, s);
}

int merge_topk(int P, const float* dist, const long long* idx, long long nq, int k_in, int k_out,
               float* out_dist, long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s) {
  return merge_topk_strided(P, dist, nq * (long long)k_in, idx, nq * (long long)k_in, nq, k_in, k_out, out_dist,
                            out_idx, ws, ws_bytes, s);
}

// ---------------------------------------------------------------------------------------
// post-validator: the reference's acceptance rule for every returned pair
// (parquet_to_format.py:351-383: isclose(1 - dot(q, base[idx]), distance / 2)), one warp per pair.
// HBM gather bound: Q*k rows of 4d bytes.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
validate_pairs_kernel(const float* __restrict__ base, long long nb, const float* __restrict__ query, long long nq,
                      int d, const long long* __restrict__ idx, const float* __restrict__ dist, int k,
                      double rtol, double atol, unsigned long long* __restrict__ mismatches) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pairs = nq * k;
  unsigned long long bad = 0;
  for (long long pidx = (long long)blockIdx.x * 8 + warp; pidx < pairs; pidx += (long long)gridDim.x * 8) {
    const long long q = pidx / k;
    const long long row = idx[pidx];
    double acc = 0.0;
    bool nonzero = false;
    if (row >= 0 && row < nb) {
      const float* qrow = query + q * d;
      const float* brow = base + row * d;
      for (int c = lane; c < d; c += 32) {
        const float qv = __ldg(qrow + c);
        nonzero |= (qv != 0.f);
        acc = fma((double)qv, (double)__ldg(brow + c), acc);
      }
    }
    acc = warp_sum(acc);
    nonzero = __any_sync(0xffffffffu, nonzero);
    if (lane == 0) {
      // zero query vectors are skipped by the reference (parquet_to_format.py:363-366)
      if (row < 0 || row >= nb) bad += 1;
      else if (nonzero) {
        const double a = 1.0 - acc, b = (double)dist[pidx] / 2.0;
        if (!(fabs(a - b) <= atol + rtol * fabs(b))) bad += 1;      // numpy.isclose(a, b)
      }
    }
  }
  if (lane == 0 && bad) atomicAdd(mismatches, bad);
}

int validate_pairs(const float* base, long long nb, const float* query, long long nq, int d, const long long* idx,
                   const float* dist, int k, double rtol, double atol, long long* out_host, cudaStream_t s) {
  unsigned long long* dcount = nullptr;
  NWK_CUDA(cudaMalloc(&dcount, 8));
  NWK_CUDA(cudaMemsetAsync(dcount, 0, 8, s));
  long long blocks = (nq * k + 7) / 8;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  validate_pairs_kernel<<<(unsigned)blocks, 256, 0, s>>>(base, nb, query, nq, d, idx, dist, k, rtol, atol, dcount);
  count_launch();
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, dcount, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(dcount);
  if (e != cudaSuccess) return fail(NWK_ERR_CUDA, "validate_pairs failed: %s", cudaGetErrorString(e));
  *out_host = (long long)h;
  return NWK_OK;
}

}  // namespace nwk
