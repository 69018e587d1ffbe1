// nwk_tc.cu -- tensor-core candidate stage + fp64 finish for exact brute-force kNN.
//
// Pipeline of one nwk_knn_f32 call (replaces knn() at cu_knn.py:243-257 of the reference):
//
//   qmax / qprep     queries -> fp16 with one global power-of-two scale (exact scaling).
//   baseprep         per base row j: |b_j|^2 in fp64, alpha_j / beta_j so that
//                    key(q,j) = alpha_j + beta_j * acc(q,j) orders rows like the requested
//                    distance (|b|^2 - 2 q.b, -q.b/|b|, -q.b); fp16 copy of the row with a
//                    per-row power-of-two scale (folded into beta_j).  HBM-bound: reads
//                    B*d*4 bytes, writes B*d*2.
//   tc_kernel<SAMPLE>  tcgen05 GEMM of every query tile against a stratified row sample of the base
//                    (one pseudo-random row per block of 32, gathered by baseprep into a compact
//                    matrix); the epilogue keeps the 32 smallest keys per query in registers
//                    -> threshold tau_q with about 1024 rows of the whole base below it.
//   tc_kernel<FILTER>  the same GEMM over the whole base; the epilogue compares each key with
//                    tau_q straight out of TMEM and appends the (rare) survivors to a
//                    per-query candidate log.  The distance matrix never reaches HBM.
//                    Tensor-pipe / L2 bound: 2*Q*B*d FLOP.
//   select_kernel    per query: radix-select the k-th smallest logged key, keep every candidate
//                    within 2*eps of it (eps = proven error bound of the fp16 operands).
//                    A query is *certified* when tau_q lies above that band (no row outside the
//                    log can belong to the true top-k); the rest go to the fp64 scan.
//   rerank_kernel    recompute the kept rows in fp64 from the fp32 inputs (HBM gather bound).
//   output_kernel    order by (fp64 distance, index), emit k, check the error bound held.
//
// Operand precision: fp16 has the 11-bit significand of TF32; with exact power-of-two scaling
// (per base row, global for the queries) every element is a normal fp16 number, so the
// candidate keys are as accurate as TF32 ones while the tensor pipe runs at twice the TF32
// rate and the operand tiles are half as large (the kernel is L2->SM bandwidth bound).  The
// TF32 variant (fp32 operands straight from the caller's buffers, no staging copy) is kept
// behind NWK_FLAG_TF32.
//
// Tile shape: 128 queries (MMA M, one TMEM lane per query) x 256 base rows (MMA N), K-blocks
// of 128 bytes (= one swizzle span: 64 fp16 / 32 fp32), 4-stage TMA ring (48 KB per stage),
// two TMEM accumulator stages of 256 columns so the epilogue of pair p overlaps the MMAs of
// pair p+1.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <vector>

#include "nwk_common.cuh"
#include "nwk_ptx.cuh"

namespace nwk {

int exact_scan(const float* base, long long nb, const float* query, int d, int k, int metric,
               long long id_offset, const int* qsel, long long nsel, float* out_dist,
               long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s);
size_t exact_scan_workspace_bytes(long long nb, long long nsel, int k);

constexpr int TQ = 128;
constexpr int TB = 256;
constexpr int STAGES = 4;
constexpr int PAIR_STAGES = 6;                  // operand ring of the CTA-pair filter kernel (32 KB per stage and CTA)
constexpr int ACC_STAGES = 2;
constexpr int TC_THREADS = 384;                // 4 control warps + 2 epilogue warpgroups
constexpr int TC_THREADS_WIDE = 640;           // ... + 4 epilogue warpgroups (direct keys on a single CTA, see epi_wg())
constexpr int EPI_WG_MAX = 2;
constexpr int STAGE_A_BYTES = TQ * 128;        // 16 KB: 128 rows x one 128-byte swizzle span
constexpr int STAGE_B_BYTES = TB * 128;        // 32 KB
constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
constexpr int AB_BYTES = ACC_STAGES * 2 * TB * 4;        // alpha + beta of the tile in each accumulator stage
constexpr int APPEND_SLOTS = 8;                          // per-thread staging of candidate records
constexpr int APPEND_BYTES = APPEND_SLOTS * TQ * EPI_WG_MAX * 8;
constexpr int MASK_BYTES = (TB / 32) * TQ * 4;           // per-thread hit masks of its column chunks
constexpr int TC_SMEM_BYTES = STAGES * STAGE_BYTES + AB_BYTES + APPEND_BYTES + MASK_BYTES + 256 + 1024;
constexpr int SAMPLE_M = 32;                   // per-query sample list length (registers)
constexpr int SAMPLE_WARM = 4;                 // warm-up: the 4 smallest of a part's first 4 chunks (128 sample rows) set the first threshold
constexpr int SAMPLE_STAGES = 3;               // operand ring of the sample pass; the fourth stage's 48 KB hold the key queues
constexpr int SAMPLE_WG = 2;                   // epilogue warpgroups of the sample pass, each with its own lists over half of the columns
constexpr int SAMPLE_QCAP = STAGE_BYTES / (SAMPLE_WG * TQ * 4);       // per-thread queue of admitted keys (shared memory)
constexpr int SAMPLE_MAX_PARTS = 8;            // parts a query tile's sample scan may be split into
constexpr int SAMPLE_MAX_LISTS = SAMPLE_MAX_PARTS * SAMPLE_WG;        // lists per query that tau_merge_kernel unites
static_assert(SAMPLE_QCAP >= 40, "the queue must take a whole 32-column chunk beyond the flush mark");
constexpr int LOG_CAP = 4096;                  // candidate-log slots per query
constexpr int LOG_TARGET = 1024;               // population the sampled threshold aims at
constexpr int RERANK_MAX = 2048;               // most rows re-ranked in fp64 per query (fp64 results live in log slots [LOG_CAP/2, LOG_CAP))
static_assert(RERANK_MAX <= LOG_CAP / 2, "re-ranked distances alias the upper half of the query's log");
constexpr int TC_MAX_K = 256;
constexpr int RETRY_MAX = 4096;                // most uncertified queries of a call that get a second tensor-core pass
constexpr unsigned int FLAG_INTERNAL_RETRY = 0x80000000u;   // the call IS that second pass (never set by callers)
#ifndef NWK_WARP_ARRIVE
#define NWK_WARP_ARRIVE 1         // 1: one tempty arrival per epilogue warp, 0: one per thread
#endif
enum { MODE_SAMPLE = 0, MODE_FILTER = 1, MODE_DUMP = 2 };
enum { ERR_PRODUCER = 101, ERR_MMA_EMPTY = 102, ERR_MMA_FULL = 103, ERR_EPI = 104 };

// operand-type traits: elements per 128-byte K-block, K per MMA instruction
template <bool F16> struct Op {
  static constexpr int BK = F16 ? 64 : 32;
  static constexpr int UMMA_K = F16 ? 16 : 8;
};

struct Scalars;
struct TcParams {
  int nKB;
  int kLast;                // MMA K-steps of the last K-block (the row pitch need not fill it; TMA zero-fills the rest)
  long long nq, nb;
  int nQT, nT;
  int G, nChunks, nUnits;   // FILTER: unit = (chunk of G base tiles, query tile)
  int t0;                   // FILTER: first base tile of this launch (the base may be filtered in parts; nT counts this part's tiles)
  int sM;                   // SAMPLE: rank of the list entry that becomes the threshold (a multiple of 4 in [8, SAMPLE_M])
  int sWarm;                // SAMPLE: 1 = start every part with the short warm-up list (large samples only)
  int nS, sChunks, sParts;  // SAMPLE: tiles of the sample matrix, chunks per tile (8), parts per query tile (<= SAMPLE_MAX_PARTS)
  const float* alpha;       // [nT*TB]   (SAMPLE mode: the sample rows' values, [nS*TB])
  const float* beta;        // [nT*TB]
  float* tau;               // [nQT*TQ]
  float* slist;             // SAMPLE: [nq][sParts][SAMPLE_M] smallest keys per query and sample part
  unsigned int* cnt;        // [nq]
  unsigned long long* log;  // [nq][LOG_CAP]
  const Scalars* sc;        // direct keys: the scales that make up kscale
  int metric;
  unsigned long long* prof; // NWK_DEBUG_PROF: per CTA, cycles the MMA issuer spent waiting for (accumulator, operands) and in total, tiles
  int dbg_na;               // NWK_DEBUG_NA = 1: a single resident query buffer even where two fit
  int dbg;                  // NWK_DEBUG_EPI (timing experiments, results are wrong): 1 = no key passes, 2 = epilogue skips the accumulators,
                            // 3 = (four-warpgroup epilogue) the accumulators are read from TMEM and dropped
  float* dump;              // MODE_DUMP: [TQ][TB]
  int dump_qt, dump_t;
  int* err;
};

struct Scalars {            // one block of device scalars, zeroed per call
  unsigned int bn2max;      // max |b|^2 (float bits)
  unsigned int qabsmax;     // max |q_i| (float bits)
  int err;
  int fail_cnt;
  unsigned int stat_max[2];
  unsigned int err_ratio;   // max |approx key - exact key| / eps over re-ranked rows (float bits)
  unsigned int babsmax;     // max finite |b_ij| (float bits): the global base scale of the direct-key filter
  unsigned int qn2max;      // max |q|^2 (float bits): the common scale of direct squared-L2 keys
  int overflow;             // direct squared-L2 keys: some base row does not fit the scale taken from the queries
  unsigned long long stat_sum[2];
};

// ---------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
qmax_kernel(const float* __restrict__ q, long long n, unsigned int* __restrict__ out) {
  // largest finite magnitude (fmaxf drops NaN, infinities are skipped); n % 4 == 0 (d % 4 == 0)
  const float INF = __int_as_float(0x7f800000);
  float m = 0.f;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n >> 2); i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(q4 + i);
    const float a0 = fabsf(v.x), a1 = fabsf(v.y), a2 = fabsf(v.z), a3 = fabsf(v.w);
    m = fmaxf(m, fmaxf(fmaxf(a0 < INF ? a0 : 0.f, a1 < INF ? a1 : 0.f), fmaxf(a2 < INF ? a2 : 0.f, a3 < INF ? a3 : 0.f)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// power-of-two exponent s such that maxabs * 2^s lies in [2^14, 2^15): every element of the
// scaled vector is then far inside fp16's range and only elements below 2^-28 * maxabs fall
// into the subnormal range.
__device__ __forceinline__ int f16_scale_exp(float maxabs) {
  if (!(maxabs > 0.f)) return 0;
  int e;
  frexpf(maxabs, &e);       // maxabs = m * 2^e, m in [0.5, 1)
  // clamped so that the combined query+row scale stays a normal fp32 power of two; data whose
  // magnitudes lie outside 2^+-75 would leave fp16's range and is caught by the error monitor
  return max(-60, min(60, 15 - e));
}

__device__ __forceinline__ float pow2f(int e);
__device__ __forceinline__ void direct_exps(int metric, const Scalars* sc, int& sq, int& sb);

// max |q|^2 over the queries (one warp per row); only direct squared-L2 keys need it
__global__ void __launch_bounds__(256)
qnorm_kernel(const float* __restrict__ q, long long nq, int d, unsigned int* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = 0.f;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < nq; r += (long long)gridDim.x * 8) {
    float a = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = q[r * d + c]; a = fmaf(v, v, a); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (a < __int_as_float(0x7f800000)) m = fmaxf(m, a);
  }
  if (lane == 0 && m > 0.f) atomicMax(out, __float_as_uint(m * 1.0001f));   // (fp32 summation slack)
}

// Direct squared-L2 keys put the -|b|^2/2 term of the key into the GEMM: the operand copies get
// L2X_COLS extra K columns, the queries hold the powers of two L2X_Q there and every base row
// three fp16 pieces h1, h2, h3 of X = 2^(2s-1) |b|^2 with h1 2^15 + h2 2^4 + h3 2^-7 = X (to 2^-29
// relative), negated, so that acc' = 2^2s (q.b - |b|^2/2) and key = -2^(1-2s) acc'.
constexpr int L2X_COLS = 8;
__device__ __forceinline__ float l2x_q(int c) { return c == 0 ? 32768.f : (c == 1 ? 16.f : (c == 2 ? 0.0078125f : 0.f)); }

__global__ void __launch_bounds__(256)
qprep_kernel(const float* __restrict__ q, long long nq, int d, int dh, int pitch, int metric, int direct,
             const Scalars* __restrict__ sc, __half* __restrict__ qh) {
  int s = f16_scale_exp(__uint_as_float(sc->qabsmax)), sb_unused;
  if (direct) direct_exps(metric, sc, s, sb_unused);
  const long long n = nq * pitch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / pitch;
    int c = (int)(i - r * pitch);
    float v = c < d ? q[r * d + c] * pow2f(s) : 0.f;
    if (c >= dh) v = (direct && metric == NWK_METRIC_SQEUCLIDEAN) ? l2x_q(c - dh) : 0.f;   // the |b|^2 columns' partners; padding
    qh[i] = __float2half_rn(v);
  }
}

// exact 2^e as a float (e in [-126, 127])
__device__ __forceinline__ float pow2f(int e) {
  e = max(-126, min(127, e));
  return __int_as_float((e + 127) << 23);
}

// Direct keys (fp16 operands, dot / cosine): every base row is scaled by the same power of two
// 2^sbg (after normalisation to unit length for cosine), so the accumulator orders the rows by
// itself: key = kscale * acc with kscale = -2^-(sq + sbg).  The filter epilogue then needs no
// per-column alpha / beta at all (their shared-memory reads bound the general epilogue).
//
// Squared L2 (any d): queries and base share one scale 2^s taken from the largest query norm
// (|q|_max 2^s in [2^12, 2^13)), because the |b|^2 term rides along as extra K columns whose
// pieces must fit fp16: X = 2^(2s-1) |b|^2 <= 2^30.9, i.e. |b| <= 11 |q|_max.  No pass over the base
// is needed for that scale; a base row that does not fit raises Scalars::overflow and the host
// repeats the call with alpha / beta keys (embeddings of one model never do).
__device__ __forceinline__ void direct_exps(int metric, const Scalars* sc, int& sq, int& sb) {
  if (metric == NWK_METRIC_SQEUCLIDEAN) {
    const float qn = sqrtf(__uint_as_float(sc->qn2max));
    int e = 13;
    if (qn > 0.f) frexpf(qn, &e);                      // qn = m 2^e, m in [0.5, 1)
    sq = sb = max(-60, min(60, 13 - e));
    return;
  }
  sq = f16_scale_exp(__uint_as_float(sc->qabsmax));
  sb = (metric == NWK_METRIC_COSINE) ? 14              // unit rows: |element| <= 1 -> <= 2^14
                                     : f16_scale_exp(__uint_as_float(sc->babsmax));
}
__device__ __forceinline__ int direct_base_exp(int direct, int metric, const Scalars* sc) {
  if (!direct) return 0;
  int sq, sb;
  direct_exps(metric, sc, sq, sb);
  return sb;
}
__device__ __forceinline__ float direct_kscale(int metric, const Scalars* sc) {
  int sq, sb;
  direct_exps(metric, sc, sq, sb);
  return metric == NWK_METRIC_SQEUCLIDEAN ? -pow2f(1 - (sq + sb)) : -pow2f(-(sq + sb));
}

// one warp per base row: |b|^2 (fp64), alpha/beta, optional fp16 copy with a per-row scale.
// The row is read once: up to ROW_CACHE float4 per lane stay in registers between the
// norm / max pass and the conversion pass (d <= 1024); longer rows re-read their tail (L1/L2).
// Stratified row sample: block i of `stride` consecutive rows contributes the row
// i*stride + hash(i) % (rows in the block).  One pseudo-random row per block keeps the sample
// representative for bases stored cluster by cluster (a tile-granular sample is not).
__host__ __device__ __forceinline__ long long sample_pick(long long blk, int stride_log2, long long nb) {
  unsigned long long z = (unsigned long long)blk * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const long long lo = blk << stride_log2;
  long long pick = lo + (long long)(z & ((1ull << stride_log2) - 1));
  if (pick >= nb) pick = lo + (long long)(z % (unsigned long long)(nb - lo));   // short last block
  return pick;
}

struct SampleOut {
  int stride_log2;          // rows per block = 2^stride_log2; < 0 = no sample wanted
  long long ns, ns_pad;     // sample rows, padded to whole tiles
  void* rows;               // [ns, dh] fp16 (F16) or [ns, d] fp32 (TF32): the operand copy of the sampled rows
  float* alpha;             // [ns_pad]
  float* beta;              // [ns_pad]
};

// Two launch modes:
//   rows     (so.stride_log2 < 0): rows [row_begin, row_end) of the padded base -> alpha / beta and the
//            fp16 operand copy.  The base may be prepared in parts, on a side stream, while the filter
//            pass already runs over the parts that are done (tc_knn).
//   sample   (so.stride_log2 >= 0): only the stratified sample rows (one per block of 2^stride_log2
//            rows) -> the compact sample matrix with its own alpha / beta, which the sample pass needs
//            before any threshold exists.  Reads 1/32 .. 1/64 of the base.
// ROW_CACHE = float4 per lane kept in registers (8: rows up to 1024 columns are read once; longer rows re-read
// their tail).  Short rows (d <= 256) take the ROW_CACHE = 2 instantiation: half the registers, twice the
// warps per SM -- a warp has one 512-byte row in flight at d = 128, so the resident warps are the HBM
// parallelism of that shape.
template <bool F16, int ROW_CACHE = 8>
__global__ void __launch_bounds__(256, ROW_CACHE <= 2 ? 6 : 1)
baseprep_kernel(const float* __restrict__ base, long long nb, long long row_begin, long long row_end, int d, int dh,
                int metric, float* __restrict__ alpha, float* __restrict__ beta, Scalars* __restrict__ sc,
                __half* __restrict__ bh, const SampleOut so, const int direct, const int pitch) {
  // pitch = elements per row of the fp16 copies: dh, or dh + L2X_COLS for direct squared-L2 keys
  const bool sample_mode = so.stride_log2 >= 0;
  if (sample_mode) {
    const float PINF = __int_as_float(0x7f800000);
    for (long long i = so.ns + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < so.ns_pad;
         i += (long long)gridDim.x * blockDim.x) {
      so.alpha[i] = PINF;
      so.beta[i] = 0.f;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float INF = __int_as_float(0x7f800000);
  int sq = F16 ? f16_scale_exp(__uint_as_float(sc->qabsmax)) : 0;
  // direct keys (F16 only): every row shares one scale, so key = kscale * acc for the whole call
  int sbg = 0;
  if (direct) direct_exps(metric, sc, sq, sbg);
  const bool l2x = direct && metric == NWK_METRIC_SQEUCLIDEAN;
  float local_max = 0.f;
  const long long n_items = sample_mode ? so.ns : row_end - row_begin;
  const int wpb = blockDim.x >> 5;             // 8 warps per CTA, or 4 when the CTAs must fit beside a filter CTA
  for (long long it = (long long)blockIdx.x * wpb + warp; it < n_items; it += (long long)gridDim.x * wpb) {
    const long long srow = sample_mode ? it : -1;
    const long long row = sample_mode ? sample_pick(it, so.stride_log2, nb) : row_begin + it;
    if (row >= nb) {
      if (lane == 0) { alpha[row] = INF; beta[row] = 0.f; }
      continue;
    }
    const float* brow = base + row * d;
    float4 cache[ROW_CACHE];
#pragma unroll
    for (int i = 0; i < ROW_CACHE; ++i) {
      const int c = lane * 4 + i * 128;
      cache[i] = c < d ? __ldg(reinterpret_cast<const float4*>(brow + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double a = 0.0;
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_CACHE; ++i) {
      // (warp-uniform: short rows skip the empty part of the cache -- its conversions and DFMAs on zeros kept
      //  the fp64 pipe busy for 8 x 128 columns whatever d was: 1.17 ms for 2.25M rows of d = 128, 1.5 TB/s)
      if (i * 128 >= d) break;
      const float4 v = cache[i];
      a = fma((double)v.x, (double)v.x, a);
      a = fma((double)v.y, (double)v.y, a);
      a = fma((double)v.z, (double)v.z, a);
      a = fma((double)v.w, (double)v.w, a);
      mx = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(mx, fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int c = lane * 4 + ROW_CACHE * 128; c < d; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(brow + c));
      a = fma((double)v.x, (double)v.x, a);
      a = fma((double)v.y, (double)v.y, a);
      a = fma((double)v.z, (double)v.z, a);
      a = fma((double)v.w, (double)v.w, a);
      mx = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(mx, fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    a = warp_sum(a);
    int sb = 0;
    bool row_ok = true;
    if (!F16 && sample_mode) {
      float* frow = reinterpret_cast<float*>(so.rows) + srow * d;
#pragma unroll
      for (int i = 0; i < ROW_CACHE; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < d) *reinterpret_cast<float4*>(frow + c) = cache[i];
      }
      for (int c = lane * 4 + ROW_CACHE * 128; c < d; c += 128)
        *reinterpret_cast<float4*>(frow + c) = __ldg(reinterpret_cast<const float4*>(brow + c));
    }
    if (F16) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (!(mx < INF)) mx = 0.f;             // inf / nan rows: leave unscaled (they never rank)
      sb = f16_scale_exp(mx);
      float scale = pow2f(sb);               // exact power of two: v * scale is exact up to fp16 rounding
      if (direct) {
        // rows that cannot rank (a nan entry, a zero vector under cosine) become zero rows flagged
        // by alpha = +inf; the filter's slow path adds alpha to the key of every hit.  (Rows with
        // infinite entries keep them: their accumulators come out +-inf / nan like their distances.)
        row_ok = (a == a) && (metric != NWK_METRIC_COSINE || a > 0.0);
        if (l2x && row_ok && !(ldexp(a, 2 * sbg - 1) <= 2.0e9)) {
          // |b|^2 does not fit the scale taken from the queries: the host repeats the call with alpha / beta keys
          row_ok = false;
          if (lane == 0) sc->overflow = 1;
        }
        sb = sbg;
        scale = pow2f(sbg);
        if (metric == NWK_METRIC_COSINE) scale = (float)(1.0 / sqrt(a)) * scale;   // unit row at 2^sbg
        if (!row_ok) scale = 0.f;
      }
      // the row's fp16 copy goes to the operand matrix (rows mode) or to the compact sample matrix
      __half* hrow = sample_mode ? reinterpret_cast<__half*>(so.rows) + srow * pitch : bh + row * pitch;
      if (l2x && lane == 0) {
        // X = 2^(2s-1) |b|^2 as three fp16 pieces against the queries' 2^15, 2^4, 2^-7 (exact residuals in fp64)
        float h1 = 0.f, h2 = 0.f, h3 = 0.f;
        if (row_ok) {
          const double X = ldexp(a, 2 * sbg - 1);
          h1 = __half2float(__float2half_rn((float)(X * (1.0 / 32768.0))));
          const double r1 = X - (double)h1 * 32768.0;
          h2 = __half2float(__float2half_rn((float)(r1 * (1.0 / 16.0))));
          const double r2 = r1 - (double)h2 * 16.0;
          h3 = __half2float(__float2half_rn((float)(r2 * 128.0)));
        }
        __half2 p01 = __floats2half2_rn(-h1, -h2);
        __half2 p23 = __floats2half2_rn(-h3, 0.f);
        uint4 pk4 = make_uint4(*reinterpret_cast<unsigned int*>(&p01), *reinterpret_cast<unsigned int*>(&p23), 0u, 0u);
        *reinterpret_cast<uint4*>(hrow + dh) = pk4;
      }
#pragma unroll
      for (int i = 0; i < ROW_CACHE; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < dh) {                        // dh % 8 == 0; columns in [d, dh) are zero
          const float4 v = cache[i];
          __half2 h01 = __floats2half2_rn(v.x * scale, v.y * scale);
          __half2 h23 = __floats2half2_rn(v.z * scale, v.w * scale);
          uint2 pk;
          pk.x = *reinterpret_cast<unsigned int*>(&h01);
          pk.y = *reinterpret_cast<unsigned int*>(&h23);
          if (!row_ok) pk = make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(hrow + c) = pk;
        }
      }
      for (int c = lane * 4 + ROW_CACHE * 128; c < dh; c += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) v = __ldg(reinterpret_cast<const float4*>(brow + c));
        __half2 h01 = __floats2half2_rn(v.x * scale, v.y * scale);
        __half2 h23 = __floats2half2_rn(v.z * scale, v.w * scale);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned int*>(&h01);
        pk.y = *reinterpret_cast<unsigned int*>(&h23);
        if (!row_ok) pk = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(hrow + c) = pk;
      }
    }
    if (lane == 0) {
      // acc = 2^(sq+sb) * (q.b) up to operand rounding: fold the scale into beta (exact)
      const float unscale = pow2f(-(sq + sb));
      float al = 0.f, be = -unscale;
      if (direct) {
        // be is the same for every row (direct_kscale); al only flags rows that must never rank
        if (l2x) be = -pow2f(1 - (sq + sb));
        if (!row_ok) { al = INF; be = 0.f; }
      } else if (metric == NWK_METRIC_SQEUCLIDEAN) { al = (float)a; be = -2.f * unscale; }
      else if (metric == NWK_METRIC_COSINE) {
        if (a > 0.0) be = (float)(-1.0 / sqrt(a)) * unscale; else { al = INF; be = 0.f; }
      }
      if (sample_mode) { so.alpha[srow] = al; so.beta[srow] = be; }
      else { alpha[row] = al; beta[row] = be; }
      local_max = fmaxf(local_max, __double2float_ru(a));
    }
  }
  if (lane == 0 && local_max > 0.f) atomicMax(&sc->bn2max, __float_as_uint(local_max));
}

// entry of rank m (1-based; m a multiple of 4 in [8, SAMPLE_M]) of a register-resident sorted list
__device__ __forceinline__ float list_rank(const float (&lst)[SAMPLE_M], int m) {
  float v = lst[SAMPLE_M - 1];
  v = m <= 28 ? lst[27] : v;
  v = m <= 24 ? lst[23] : v;
  v = m <= 20 ? lst[19] : v;
  v = m <= 16 ? lst[15] : v;
  v = m <= 12 ? lst[11] : v;
  v = m <= 8 ? lst[7] : v;
  return v;
}

// tau_q = m-th smallest key (m <= SAMPLE_M) of the union of the query's per-part lists (one warp per
// query): bitwise search for the largest ordered key t with fewer than m list entries
// below it.  Lists that were never written hold +inf.
__global__ void __launch_bounds__(256)
tau_merge_kernel(const float* __restrict__ slist, long long nq, int parts, int m, float* __restrict__ tau) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * 8 + warp;
  if (q >= nq) return;
  uint32_t v[SAMPLE_MAX_LISTS];
#pragma unroll
  for (int i = 0; i < SAMPLE_MAX_LISTS; ++i)
    v[i] = i < parts ? f2ord(slist[(q * parts + i) * SAMPLE_M + lane]) : 0xffffffffu;
  uint32_t ans = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t trial = ans | (1u << bit);
    int below = 0;
#pragma unroll
    for (int i = 0; i < SAMPLE_MAX_LISTS; ++i) below += (v[i] < trial) ? 1 : 0;
    below = __reduce_add_sync(0xffffffffu, below);
    if (below < m) ans = trial;
  }
  if (lane == 0) tau[q] = ord2f(ans);
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

// ---------------------------------------------------------------------------------------
// schedule shared by the three roles of tc_kernel: the sequence of (query tile, base tile)
// pairs this CTA walks, in an order that keeps concurrently running CTAs on the same base
// chunk (so a base tile is fetched from HBM once and served from L2 to the other CTAs).
// ---------------------------------------------------------------------------------------
template <int MODE, bool CTA2 = false>
struct Sched {
  int nUnits, nQT, nT, G, nS, sParts, dump_qt, dump_t, t0;
  int u, i, len, qt, c;
  int step;                 // workers walking the unit list: CTAs, or CTA pairs (both CTAs of a pair walk the same units)
  __device__ explicit Sched(const TcParams& p)
      : nUnits(p.nUnits), nQT(p.nQT), nT(p.nT), G(p.G), nS(p.nS), sParts(p.sParts), dump_qt(p.dump_qt), dump_t(p.dump_t),
        t0(p.t0), i(0), len(0), qt(0), c(0) {
    step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    u = (CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) - step;
  }
  // advance to the next pair; returns false when the CTA is done
  __device__ __forceinline__ bool next(int& out_qt, int& out_t, bool& first, bool& last) {
    if (i >= len) {
      if (MODE == MODE_FILTER) {
        u += step;
        if (u >= nUnits) return false;
        c = u / nQT;
        qt = u - c * nQT;
        len = min(G, nT - c * G);
      } else if (MODE == MODE_SAMPLE) {
        // unit = (query tile, part of the sample tiles); parts > 1 only when there are far fewer
        // query tiles than SMs (every part pays the warm-up of its own list)
        do {
          u += gridDim.x;
          if (u >= sParts * nQT) return false;
          qt = u / sParts;
          const int part = u - qt * sParts;
          c = (int)(((long long)part * nS) / sParts);
          len = (int)(((long long)(part + 1) * nS) / sParts) - c;
        } while (len == 0);
      } else {
        u += gridDim.x;
        if (u >= 1) return false;
        qt = dump_qt;
        len = 1;
      }
      i = 0;
    }
    out_qt = qt;
    // (all CTAs on a chunk walk its tiles in the same order on purpose: requests for a tile that is
    //  still on its way from HBM merge in L2; rotating the order per query tile was measured:
    //  +2 % filter time at d = 768, +6 % at d = 1536)
    if (MODE == MODE_FILTER) out_t = t0 + c * G + i;
    else if (MODE == MODE_SAMPLE) out_t = c + i;
    else out_t = dump_t;
    first = (i == 0);
    last = (i == len - 1);
    ++i;
    return true;
  }
};

// CTA2 = the filter pass on CTA pairs (cta_group::2): the pair owns a 256-query x 256-row tile,
// each CTA stages its own 128 query rows and 128 of the 256 base rows (32 KB per K-block instead
// of 48 KB, so six pipeline stages instead of four and a third less L2 -> SM traffic per FLOP);
// the leader CTA (cluster rank 0) issues one M=256 MMA for both; each CTA's TMEM receives its
// 128 queries x 256 rows and runs the same epilogue as the single-CTA kernel.  TcParams.nQT then
// counts query-tile PAIRS.
// ARES = the query tile stays resident in shared memory for a whole unit (up to 16 base tiles)
// instead of being re-fetched with every base tile: a third less L2 -> SM traffic per FLOP.  The
// single-CTA kernel is bound by exactly that traffic (measured with the epilogue switched off:
// 96 KB per 128 x 256 x 128 tile arrive in 1640 clk, the MMAs need 1024).  Short rows only: the
// tile's whole K extent (p.nKB x 16 KB, at most 96 KB) must fit beside a ring of base stages.
// Epilogue warpgroups of a variant.  Direct keys on a single CTA are the short-row kernels, whose
// epilogue is bound by the latency of its slow path (1.9 hit groups per warp and tile at d = 128,
// ~600 clk each, issue slots 34 % busy with two warps per scheduler): four warpgroups = four warps
// per scheduler, each over a quarter of the columns (60 registers per thread in the current build; 102 is the cap).
template <int MODE, bool CTA2, bool DIRECT>
__host__ __device__ constexpr int epi_wg() {
  return MODE == MODE_DUMP ? 1 : ((MODE == MODE_FILTER && DIRECT && !CTA2) ? 4 : 2);
}

template <int MODE, bool F16, bool CTA2 = false, bool DIRECT = false, bool ARES = false>
__global__ void __launch_bounds__(128 + 128 * epi_wg<MODE, CTA2, DIRECT>(), 1)
tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_b,
          const TcParams p) {
  static_assert(!CTA2 || (MODE == MODE_FILTER && F16), "CTA pairs are built for the fp16 filter pass");
  static_assert(!DIRECT || (MODE == MODE_FILTER && F16), "direct keys are built for the fp16 filter pass");
  static_assert(!ARES || (MODE == MODE_FILTER && F16 && !CTA2), "resident query tiles are built for the single-CTA fp16 filter pass");
  constexpr int BK = Op<F16>::BK;
  constexpr int UMMA_K = Op<F16>::UMMA_K;
  constexpr int NST = CTA2 ? PAIR_STAGES : (MODE == MODE_SAMPLE ? SAMPLE_STAGES : STAGES);
  constexpr int SB_BYTES = CTA2 ? STAGE_B_BYTES / 2 : STAGE_B_BYTES;   // base rows staged by this CTA
  constexpr int ST_BYTES = STAGE_A_BYTES + SB_BYTES;
  static_assert(NST * ST_BYTES == STAGES * STAGE_BYTES || MODE == MODE_SAMPLE, "the filter variants use the same 192 KB operand ring");
  using SchedT = Sched<MODE, CTA2>;
  const uint32_t rank = CTA2 ? ptx::cluster_ctarank() : 0u;
  // the filter pass splits the 256 accumulator columns over two epilogue warpgroups; the sample
  // pass keeps one (its per-query list lives in one thread's registers)
  constexpr int EPI_WG = epi_wg<MODE, CTA2, DIRECT>();
  // per-thread record staging: EPI_WG_MAX warpgroups x APPEND_SLOTS in s_app; the four direct-key
  // warpgroups have no alpha / beta, so theirs starts at s_ab and also takes the unused mask block
  constexpr int ASLOTS = EPI_WG > EPI_WG_MAX ? (AB_BYTES + APPEND_BYTES + MASK_BYTES) / (TQ * EPI_WG * 8) : APPEND_SLOTS;
  static_assert(ASLOTS >= 4, "record staging");
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128B-swizzled operand tiles
  // (offset arithmetic on the __shared__ pointer keeps its address space: LDS/STS, not generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_ab = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);              // [ACC_STAGES][2][TB]
  unsigned long long* s_app = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES + AB_BYTES);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + AB_BYTES + APPEND_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + AB_BYTES + APPEND_BYTES + MASK_BYTES);
  constexpr int NB = PAIR_STAGES;                  // barrier slots per ring (the longest ring any variant uses)
  uint64_t* full_bar = bars;                       // [NB]   (CTA2: only the leader's are used)
  uint64_t* empty_bar = bars + NB;                 // [NB]
  uint64_t* tfull_bar = bars + 2 * NB;             // [ACC_STAGES]
  uint64_t* tempty_bar = bars + 2 * NB + ACC_STAGES;           // (CTA2: the leader's, both CTAs' epilogues arrive)
  uint64_t* abfull_bar = bars + 2 * NB + 2 * ACC_STAGES;       // [ACC_STAGES]
  uint64_t* abempty_bar = bars + 2 * NB + 3 * ACC_STAGES;      // [ACC_STAGES] (CTA2 only)
  uint64_t* afull_bar = bars + 2 * NB + 4 * ACC_STAGES;        // [2] (ARES: resident query tile buffers)
  uint64_t* aempty_bar = bars + 2 * NB + 4 * ACC_STAGES + 2;   // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NB + 4 * ACC_STAGES + 4);
  static_assert((2 * NB + 4 * ACC_STAGES + 4) * 8 + 4 <= 256, "barrier block");
  static_assert(NST <= NB, "barrier block");
  // ARES geometry: [nA query-tile buffers of a_bytes][nsb base stages of 32 KB] inside the same 192 KB
  const int a_bytes = ARES ? p.nKB * STAGE_A_BYTES : 0;
  const int nA = ARES ? ((2 * a_bytes + 4 * STAGE_B_BYTES <= STAGES * STAGE_BYTES && p.dbg_na != 1) ? 2 : 1) : 0;
  const int nsb = ARES ? min(NB, (STAGES * STAGE_BYTES - nA * a_bytes) / STAGE_B_BYTES) : NST;
  unsigned char* const b_ring = smem + nA * a_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NB; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], CTA2 ? 2 * 4 * EPI_WG : (NWK_WARP_ARRIVE ? 4 : 128) * EPI_WG);
      ptx::mbar_init(&abfull_bar[a], 1);
      ptx::mbar_init(&abempty_bar[a], 4 * EPI_WG);
      ptx::mbar_init(&afull_bar[a], 1);
      ptx::mbar_init(&aempty_bar[a], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) { ptx::tmem_alloc_pair(s_tmem, ACC_STAGES * TB); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(s_tmem, ACC_STAGES * TB); ptx::tmem_relinquish(); }      // 512 columns
  }
  ptx::tc_fence_before();
  if (CTA2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // The three control roles run as whole, converged warps: every lane walks the schedule and polls
  // the barriers, and only the instructions that start asynchronous work sit under elect.sync.
  // (With the role under `if (lane == 0)` the compiler wraps every tcgen05 / TMA instruction into
  // an election loop and keeps the loop state in vector registers; measured on the MMA issuer:
  // 560 clk to issue 8 MMAs + 2 commits and 1550 clk per trip of a 1024-clk tile.)
  if (warp == 0) {
    // ===== TMA producer =====
    SchedT sch(p);
    int qt, t; bool first, last;
    int stage = 0; uint32_t phase = 0;
    bool ok = true;
    int abuf = 0; uint32_t aphase = 0;
    while (ok && sch.next(qt, t, first, last)) {
      if (ARES) {
        if (first) {
          // a new unit: its query tile goes into the next resident buffer (free once the MMAs
          // of the unit that used it last have retired)
          if (!ptx::mbar_wait(&aempty_bar[abuf], aphase ^ 1, p.err, ERR_PRODUCER)) { ok = false; break; }
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&afull_bar[abuf], a_bytes);
            for (int kb = 0; kb < p.nKB; ++kb)
              ptx::tma_load_2d(smem + abuf * a_bytes + kb * STAGE_A_BYTES, &tmap_q, &afull_bar[abuf], kb * BK, qt * TQ);
          }
          __syncwarp();
          if (++abuf == nA) { abuf = 0; aphase ^= 1; }
        }
        for (int kb = 0; kb < p.nKB; ++kb) {
          if (!ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err, ERR_PRODUCER)) { ok = false; break; }
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_B_BYTES);
            ptx::tma_load_2d(b_ring + stage * STAGE_B_BYTES, &tmap_b, &full_bar[stage], kb * BK, t * TB);
          }
          __syncwarp();
          if (++stage == nsb) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      for (int kb = 0; kb < p.nKB; ++kb) {
        if (!ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err, ERR_PRODUCER)) { ok = false; break; }
        unsigned char* sa = smem + stage * ST_BYTES;
        if (ptx::elect_one()) {
          if (CTA2) {
            // both CTAs' loads are credited to the leader's barrier, which expects the bytes of both
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * ST_BYTES);
            ptx::tma_load_2d_pair(sa, &tmap_q, &full_bar[stage], kb * BK, (qt * 2 + (int)rank) * TQ);
            ptx::tma_load_2d_pair(sa + STAGE_A_BYTES, &tmap_b, &full_bar[stage], kb * BK, t * TB + (int)rank * (TB / 2));
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], ST_BYTES);
            ptx::tma_load_2d(sa, &tmap_q, &full_bar[stage], kb * BK, qt * TQ);
            ptx::tma_load_2d(sa + STAGE_A_BYTES, &tmap_b, &full_bar[stage], kb * BK, t * TB);
          }
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one warp; of the leader CTA when the tile belongs to a CTA pair) =====
    if (rank == 0) {
      SchedT sch(p);
      int qt, t; bool first, last;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int abuf = 0; uint32_t aphase = 0;
      const uint32_t idesc = F16 ? ptx::umma_idesc_f16(CTA2 ? 2 * TQ : TQ, TB) : ptx::umma_idesc_tf32(TQ, TB);
      bool ok = true;
      long long w_acc = 0, w_op = 0, w_issue = 0, n_tiles = 0;
      const long long t_begin = clock64();
      while (ok && sch.next(qt, t, first, last)) {
        long long tw = p.prof ? clock64() : 0;
        if (!ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err, ERR_MMA_EMPTY)) { ok = false; break; }
        if (p.prof) { w_acc += clock64() - tw; ++n_tiles; }
        ptx::tc_fence_after();
        if (MODE != MODE_DUMP && !CTA2 && !DIRECT) {
          // the epilogue has released this stage: refill its alpha / beta tile (2 x 1 KB)
          if (ptx::elect_one()) {
            float* ab = s_ab + acc * 2 * TB;
            ptx::mbar_arrive_expect_tx(&abfull_bar[acc], 2 * TB * 4);
            ptx::bulk_load_1d(ab, p.alpha + (long long)t * TB, TB * 4, &abfull_bar[acc]);
            ptx::bulk_load_1d(ab + TB, p.beta + (long long)t * TB, TB * 4, &abfull_bar[acc]);
          }
          __syncwarp();
        }
        const uint32_t d_tmem = tmem_base + acc * TB;
        if (ARES && first) {
          if (!ptx::mbar_wait(&afull_bar[abuf], aphase, p.err, ERR_MMA_FULL)) { ok = false; break; }
        }
        for (int kb = 0; kb < p.nKB; ++kb) {
          tw = p.prof ? clock64() : 0;
          if (!ptx::mbar_wait(&full_bar[stage], phase, p.err, ERR_MMA_FULL)) { ok = false; break; }
          if (p.prof) w_op += clock64() - tw;
          ptx::tc_fence_after();
          tw = p.prof ? clock64() : 0;
          if (ptx::elect_one()) {
            const uint32_t sa = ARES ? ptx::smem_u32(smem + abuf * a_bytes + kb * STAGE_A_BYTES)
                                     : ptx::smem_u32(smem + stage * ST_BYTES);
            const uint64_t da = ptx::umma_desc_sw128(sa);
            const uint64_t db = ptx::umma_desc_sw128(ARES ? ptx::smem_u32(b_ring + stage * STAGE_B_BYTES) : sa + STAGE_A_BYTES);
            const int nk = (kb == p.nKB - 1) ? p.kLast : BK / UMMA_K;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k >= nk) break;
              // one MMA consumes 32 bytes of K inside the 128 B swizzle span: +2 in 16-byte units
              if (CTA2) ptx::mma_f16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else if (F16) ptx::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else ptx::mma_tf32_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if (CTA2) ptx::tc_commit_pair(&empty_bar[stage], 3); else ptx::tc_commit(&empty_bar[stage]);
            // last K-block: accumulator complete -> epilogue (of both CTAs of a pair)
            if (kb == p.nKB - 1) {
              if (CTA2) ptx::tc_commit_pair(&tfull_bar[acc], 3); else ptx::tc_commit(&tfull_bar[acc]);
              // the unit's MMAs are all issued: its query tile buffer is free once they retire
              if (ARES && last) ptx::tc_commit(&aempty_bar[abuf]);
            }
          }
          __syncwarp();
          if (p.prof) w_issue += clock64() - tw;
          if (++stage == nsb) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        if (ARES && last) { if (++abuf == nA) { abuf = 0; aphase ^= 1; } }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      if (p.prof && lane == 0) {
        unsigned long long* o = p.prof + (size_t)blockIdx.x * 4;
        o[0] = (unsigned long long)w_acc; o[1] = (unsigned long long)w_op | ((unsigned long long)w_issue << 32);
        o[2] = (unsigned long long)(clock64() - t_begin); o[3] = (unsigned long long)n_tiles;
      }
    }
  } else if (CTA2 && !DIRECT && warp == 3) {
    // ===== alpha / beta loader of a CTA pair: each CTA refills its own copy of the tile's
    // 256 alpha / beta values once its epilogue warps have released the accumulator stage =====
    SchedT sch(p);
    int qt, t; bool first, last;
    int acc = 0; uint32_t acc_phase = 0;
    while (sch.next(qt, t, first, last)) {
      if (!ptx::mbar_wait(&abempty_bar[acc], acc_phase ^ 1, p.err, ERR_MMA_EMPTY)) break;
      if (ptx::elect_one()) {
        float* ab = s_ab + acc * 2 * TB;
        ptx::mbar_arrive_expect_tx(&abfull_bar[acc], 2 * TB * 4);
        ptx::bulk_load_1d(ab, p.alpha + (long long)t * TB, TB * 4, &abfull_bar[acc]);
        ptx::bulk_load_1d(ab + TB, p.beta + (long long)t * TB, TB * 4, &abfull_bar[acc]);
      }
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (MODE == MODE_SAMPLE && warp >= 4 && warp < 4 + 4 * SAMPLE_WG) {
    // ===== sample epilogue: two warpgroups, each over half of every tile's columns;
    // thread <-> TMEM lane <-> query row of the tile.  Each thread keeps the SAMPLE_M smallest
    // keys of its half of its part of the sample in registers (tau_merge_kernel unites the lists).
    // Keys below the running threshold are only *queued* (one predicated shared-memory store
    // per element, no branches in the unrolled loop); the queue is merged into the list when
    // some lane's queue fills up, so a hit in one lane does not stall the other 31.  A part
    // starts with a cheap warm-up: the SAMPLE_WARM smallest keys of its first 128 rows give the
    // first threshold (about 3 % of the rows pass it), instead of 1000 full-length insertions.
    // Any threshold yields an exact result; the sample only steers the candidate count. =====
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    constexpr int QSTRIDE = SAMPLE_WG * TQ;
    constexpr int SCPW = (TB / 32) / SAMPLE_WG;      // 32-column chunks per warpgroup
    // the key queues live in the 48 KB the sample pass's shorter operand ring leaves free
    float* s_queue = reinterpret_cast<float*>(smem + SAMPLE_STAGES * STAGE_BYTES) + wg * TQ + row_in_tile;   // slot i at s_queue[i * QSTRIDE]
    Sched<MODE> sch(p);
    int qt, t; bool first, last;
    int acc = 0; uint32_t acc_phase = 0;
    const float INF = __int_as_float(0x7f800000);
    float tau = INF;
    long long q = 0;
    float lst[SAMPLE_M];
    int cnt = 0, warm = 0;
    bool ok = true;
    while (ok && sch.next(qt, t, first, last)) {
      if (first) {
        q = (long long)qt * TQ + row_in_tile;
        cnt = 0;
        // without the warm-up every key is admitted until the list is full
        warm = p.sWarm ? 0 : SAMPLE_WARM;
        tau = (p.sWarm || q < p.nq) ? INF : -INF;
#pragma unroll
        for (int j = 0; j < SAMPLE_M; ++j) lst[j] = INF;
      }
      if (!ptx::mbar_wait(&abfull_bar[acc], acc_phase, p.err, ERR_EPI)) { ok = false; break; }
      if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err, ERR_EPI)) { ok = false; break; }
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TB;
      const float4* al4 = reinterpret_cast<const float4*>(s_ab + acc * 2 * TB);
      const float4* be4 = al4 + TB / 4;

      // merge the queued keys into the list (every lane walks its own queue; +inf is a no-op)
      auto flush = [&]() {
        const int maxc = __reduce_max_sync(0xffffffffu, cnt);
        for (int i = 0; i < maxc; ++i) {
          const float key = (i < cnt) ? s_queue[i * QSTRIDE] : INF;
#pragma unroll
          for (int m = SAMPLE_M - 1; m > 0; --m) lst[m] = fmaxf(lst[m - 1], fminf(lst[m], key));
          lst[0] = fminf(lst[0], key);
        }
        cnt = 0;
        tau = fminf(tau, list_rank(lst, p.sM));
      };
      auto chunk = [&](const uint32_t (&cur)[32], int ch) {
        if (warm < SAMPLE_WARM) {                           // warp-uniform
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 al = al4[ch * 8 + j4];
            const float4 be = be4[ch * 8 + j4];
            const float k4[4] = {fmaf(be.x, __uint_as_float(cur[4 * j4 + 0]), al.x), fmaf(be.y, __uint_as_float(cur[4 * j4 + 1]), al.y),
                                 fmaf(be.z, __uint_as_float(cur[4 * j4 + 2]), al.z), fmaf(be.w, __uint_as_float(cur[4 * j4 + 3]), al.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#pragma unroll
              for (int m = SAMPLE_WARM - 1; m > 0; --m) lst[m] = fmaxf(lst[m - 1], fminf(lst[m], k4[e]));
              lst[0] = fminf(lst[0], k4[e]);
            }
          }
          if (++warm == SAMPLE_WARM) tau = (q < p.nq) ? lst[SAMPLE_WARM - 1] : -INF;
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 al = al4[ch * 8 + j4];
            const float4 be = be4[ch * 8 + j4];
            const float k4[4] = {fmaf(be.x, __uint_as_float(cur[4 * j4 + 0]), al.x), fmaf(be.y, __uint_as_float(cur[4 * j4 + 1]), al.y),
                                 fmaf(be.z, __uint_as_float(cur[4 * j4 + 2]), al.z), fmaf(be.w, __uint_as_float(cur[4 * j4 + 3]), al.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k4[e] < tau) { s_queue[cnt * QSTRIDE] = k4[e]; ++cnt; }
          }
          if (__any_sync(0xffffffffu, cnt > SAMPLE_QCAP - 32)) flush();
        }
      };

      uint32_t va[32], vb[32];
      const int c0 = wg * SCPW;                      // first chunk of this warpgroup
      ptx::tmem_ld_32x32(taddr + c0 * 32, va);
#pragma unroll 1
      for (int cp = c0; cp < c0 + SCPW; cp += 2) {
        ptx::tmem_wait_ld();
        ptx::tmem_ld_32x32(taddr + (cp + 1) * 32, vb);
        chunk(va, cp);
        ptx::tmem_wait_ld();
        if (cp + 2 < c0 + SCPW) ptx::tmem_ld_32x32(taddr + (cp + 2) * 32, va);
        chunk(vb, cp + 1);
      }
      ptx::tmem_wait_ld();
      __syncwarp();
      ptx::tc_fence_before();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);     // one arrival per epilogue warp
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      if (last) {
        flush();
        if (q < p.nq) {
          // this CTA covered one part of the query tile's sample: publish its list, tau_merge_kernel
          // takes the SAMPLE_M-th smallest of the union
          float* dst = p.slist + (q * (sch.sParts * SAMPLE_WG) + (sch.u % sch.sParts) * SAMPLE_WG + wg) * SAMPLE_M;
#pragma unroll
          for (int j = 0; j < SAMPLE_M; ++j) dst[j] = lst[j];
        }
      }
    }
  } else if (MODE != MODE_SAMPLE && warp >= 4 && warp < 4 + 4 * EPI_WG) {
    // ===== epilogue: EPI_WG warpgroups of 128 threads; thread <-> TMEM lane <-> query row of the
    // tile; warpgroup g owns the column chunks [g*CPW, (g+1)*CPW) of every accumulator tile =====
    constexpr int CPW = (TB / 32) / EPI_WG;          // 32-column chunks per warpgroup
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int etid = wg * TQ + row_in_tile;          // index of this thread's private smem slots
    SchedT sch(p);
    int qt, t; bool first, last;
    int acc = 0; uint32_t acc_phase = 0;
    const float INF = __int_as_float(0x7f800000);
    float tau = -INF;
    float thr = INF;          // DIRECT: acc > thr  <=>  kscale * acc < tau  (kscale = -2^m: exact)
    const float kscale = DIRECT ? direct_kscale(p.metric, p.sc) : 0.f;
    long long q = 0;
    int nst = 0;              // FILTER: records staged in this thread's shared-memory slots
    bool ok = true;
    while (ok && sch.next(qt, t, first, last)) {
      if (first) {
        q = (long long)(CTA2 ? qt * 2 + (int)rank : qt) * TQ + row_in_tile;
        if (MODE == MODE_FILTER) tau = (q < p.nq && p.dbg != 1) ? p.tau[q] : -INF;
        if (DIRECT) thr = tau / kscale;
      }
      if (MODE != MODE_DUMP && !DIRECT) {
        if (!ptx::mbar_wait(&abfull_bar[acc], acc_phase, p.err, ERR_EPI)) { ok = false; break; }
      }
      if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err, ERR_EPI)) { ok = false; break; }
      ptx::tc_fence_after();
      const int c0 = wg * CPW;                       // first chunk of this warpgroup
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TB;
      const float* s_al = s_ab + acc * 2 * TB;
      const float* s_be = s_al + TB;
      const float4* al4 = reinterpret_cast<const float4*>(s_al);
      const float4* be4 = reinterpret_cast<const float4*>(s_be);
      const long long row0 = (long long)t * TB;

      // ---- fast path: per group of 8 columns one running minimum of the keys (FMNMX3) and one
      // compare; a lane only records WHICH groups hold a key below its threshold (bit ci*4+g of
      // `hits`).  General keys: alpha_j + beta_j * acc (one FFMA per element, alpha / beta read as
      // broadcast LDS.128 -- those reads are what bounds this variant).  Direct keys: the
      // accumulator itself against thr = tau / kscale, no shared-memory traffic at all. ----
      uint32_t hits = 0;
      if (EPI_WG > EPI_WG_MAX) {
        // four warpgroups (direct keys): 64 columns per warp as four 16-column loads, two in flight,
        // (60 registers per thread in the current build: the cap of 640 threads is 102)
        if (p.dbg != 2) {
          uint32_t wa[16], wb[16];
          ptx::tmem_ld_32x16(taddr + c0 * 32, wa);
#pragma unroll
          for (int h = 0; h < CPW * 2; ++h) {
            ptx::tmem_wait_ld();
            uint32_t(&cur)[16] = (h & 1) ? wb : wa;
            uint32_t(&nxt)[16] = (h & 1) ? wa : wb;
            if (h + 1 < CPW * 2) ptx::tmem_ld_32x16(taddr + c0 * 32 + (h + 1) * 16, nxt);
            if (p.dbg == 3) continue;                     // (timing experiment: the tile is drained from TMEM, nothing else)
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float m = fmaxf(__uint_as_float(cur[8 * g]), __uint_as_float(cur[8 * g + 1]));
#pragma unroll
              for (int j = 2; j < 8; j += 2)
                m = fmaxf(m, fmaxf(__uint_as_float(cur[8 * g + j]), __uint_as_float(cur[8 * g + j + 1])));
              if (m > thr) hits |= 1u << (h * 2 + g);     // bit = group of 8 columns, counted from this warp's first column
            }
          }
          ptx::tmem_wait_ld();
        }
      } else {
      uint32_t va[32], vb[32];
      if (p.dbg != 2) {
      ptx::tmem_ld_32x32(taddr + c0 * 32, va);
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int ch = c0 + ci;
        ptx::tmem_wait_ld();
        uint32_t(&cur)[32] = (ci & 1) ? vb : va;
        uint32_t(&nxt)[32] = (ci & 1) ? va : vb;
        if (ci + 1 < CPW) ptx::tmem_ld_32x32(taddr + (ch + 1) * 32, nxt);
        if (MODE == MODE_DUMP) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            p.dump[(long long)row_in_tile * TB + ch * 32 + j] = __uint_as_float(cur[j]);
        } else if (DIRECT) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float m = fmaxf(__uint_as_float(cur[8 * g]), __uint_as_float(cur[8 * g + 1]));
#pragma unroll
            for (int j = 2; j < 8; j += 2)
              m = fmaxf(m, fmaxf(__uint_as_float(cur[8 * g + j]), __uint_as_float(cur[8 * g + j + 1])));
            if (m > thr) hits |= 1u << (ci * 4 + g);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float m = INF;
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              const float4 al = al4[ch * 8 + g * 2 + j4];           // shared-memory broadcast reads
              const float4 be = be4[ch * 8 + g * 2 + j4];
              const int e = 8 * g + 4 * j4;
              m = fminf(m, fminf(fmaf(be.x, __uint_as_float(cur[e + 0]), al.x), fmaf(be.y, __uint_as_float(cur[e + 1]), al.y)));
              m = fminf(m, fminf(fmaf(be.z, __uint_as_float(cur[e + 2]), al.z), fmaf(be.w, __uint_as_float(cur[e + 3]), al.w)));
            }
            if (m < tau) hits |= 1u << (ci * 4 + g);
          }
        }
      }
      ptx::tmem_wait_ld();
      }
      }

      // ---- slow path (rare): groups where some lane of the warp has a hit.  One rolled loop:
      // re-read the group's 8 columns from TMEM, recompute the keys (bit-identical to the fast
      // path) and stage the records of the lanes that hit.  Every branch in it is warp-uniform
      // (votes) and the per-lane work is predicated: divergent branches and their reconvergence
      // cost more here than the arithmetic (measured: 30 % of the loop's stall samples).
      // (Measured alternatives at d = 128: doing this work inside the fast path, while the values
      // are still in registers, behind one vote per group: +18 % filter time; four groups per
      // trip with their loads and compare chains interleaved: +9 %; releasing the accumulator
      // stage first and working on hit groups stashed in registers afterwards: +2..4 % -- the
      // warps' total work per tile, not the slowest warp, is what the tile waits for.  Round 2,
      // four warpgroups, experiment build on one box: recording the hits from the fast path's
      // registers under a divergent branch per group, so that no warp makes this second trip to
      // TMEM: 16.8 -> 18.9 ms.
      // What the loop costs there: 16.8 ms with it, 15.0 with no key passing, 14.1 with the
      // accumulators read and dropped, 14.2 with the epilogue not touching them -- profiles/r02_summary.md.) ----
      if (MODE == MODE_FILTER) {
        uint32_t um = __reduce_or_sync(0xffffffffu, hits);
        unsigned long long* const slots = (EPI_WG > EPI_WG_MAX ? reinterpret_cast<unsigned long long*>(s_ab) : s_app) + etid;
        while (um) {                                   // warp-uniform
          const int b = __ffs(um) - 1;
          um &= um - 1;
          const int colb = c0 * 32 + b * 8;            // first column of the group inside the tile
          uint32_t v[8];
          ptx::tmem_ld_32x8(taddr + colb, v);
          ptx::tmem_wait_ld();
          float key[8];
          uint32_t m8 = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (DIRECT) {
              key[j] = __uint_as_float(v[j]);
              if (key[j] > thr) m8 |= 1u << j;
            } else {
              key[j] = fmaf(s_be[colb + j], __uint_as_float(v[j]), s_al[colb + j]);
              if (key[j] < tau) m8 |= 1u << j;
            }
          }
          do {                                         // one trip unless a lane hit twice in the group
            const bool has = m8 != 0;
            const int j = has ? __ffs(m8) - 1 : 0;
            m8 &= m8 - 1;                              // (0 stays 0)
            float kj = key[0];
#pragma unroll
            for (int jj = 1; jj < 8; ++jj) kj = (j == jj) ? key[jj] : kj;
            if (DIRECT) kj *= kscale;
            const long long row = row0 + colb + j;
            // direct keys carry no alpha: the zero rows of the tile padding are dropped here, the
            // (data-dependent, rare) rows flagged by alpha = +inf when select_kernel reads the log
            const bool take = has && (!DIRECT || row < p.nb);
            if (take)
              slots[nst * (TQ * EPI_WG)] = (static_cast<unsigned long long>(f2ord(kj)) << 32) |
                                           static_cast<unsigned long long>(row);
            nst += take ? 1 : 0;
            if (__any_sync(0xffffffffu, nst == ASLOTS)) {
              if (nst == ASLOTS) {
                // one global atomic reserves log space for the lane's batch of staged records
                const unsigned int base = atomicAdd(p.cnt + q, (unsigned int)ASLOTS);
#pragma unroll
                for (int i = 0; i < ASLOTS; ++i)
                  if (base + i < LOG_CAP) p.log[q * LOG_CAP + base + i] = slots[i * (TQ * EPI_WG)];
                nst = 0;
              }
            }
          } while (__any_sync(0xffffffffu, m8 != 0));
        }
      }
      __syncwarp();
      ptx::tc_fence_before();
      if (CTA2) {
        // the accumulator stage belongs to the pair's MMA issuer (leader CTA); alpha / beta are local
        if (lane == 0) { ptx::mbar_arrive_cluster(&tempty_bar[acc], 0); if (!DIRECT) ptx::mbar_arrive(&abempty_bar[acc]); }
      } else {
#if NWK_WARP_ARRIVE
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);     // one arrival per epilogue warp
#else
        ptx::mbar_arrive(&tempty_bar[acc]);
#endif
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      if (MODE == MODE_FILTER && last && nst > 0) {
        const unsigned int base = atomicAdd(p.cnt + q, (unsigned int)nst);
        for (int j = 0; j < nst; ++j)
          if (base + j < LOG_CAP)
            p.log[q * LOG_CAP + base + j] =
                (EPI_WG > EPI_WG_MAX ? reinterpret_cast<unsigned long long*>(s_ab) : s_app)[j * (TQ * EPI_WG) + etid];
        nst = 0;
      }
    }
  }

  ptx::tc_fence_before();
  if (CTA2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CTA2) ptx::tmem_dealloc_pair(tmem_base, ACC_STAGES * TB);
    else ptx::tmem_dealloc(tmem_base, ACC_STAGES * TB);
  }
}

// ---------------------------------------------------------------------------------------
// finish: select within the proven error band, re-rank in fp64, emit top-k, certify
// ---------------------------------------------------------------------------------------
struct FinParams {
  const float* base; const float* query;
  long long nb, nq; int d, k, metric;
  long long id_offset;
  const float* tau; const unsigned int* cnt;
  unsigned long long* log;     // [nq][LOG_CAP]; select rewrites [0,R) sorted, rerank writes fp64 at [LOG_CAP/2, LOG_CAP/2+R)
  int* rr;                     // [nq] rows to re-rank (approximate key <= a_k + eps), -1 = uncertified
  int* rr2;                    // [nq] rows in the outer band (a_k + eps, a_k + 2 eps], parked at the end of the log
  float* epsq;                 // [nq] proven key error bound
  double* qn2;                 // [nq] |q|^2
  Scalars* sc;
  float c_rel;                 // relative error constant of the operand format
  float c_acc;                 // its accumulation part alone: 4 ulp(fp32) per MMA step
  int f16;                     // fp16 operands: adds the subnormal-quantisation term
  int direct;                  // direct keys: one global base scale (base-side quantisation term, cosine pre-normalisation)
  const float* alpha;          // direct keys: +inf marks base rows that must never rank
  const unsigned char* bflag;  // [nq] or null: 1 = the caller's distance bound lowered tau_q (nwk_topk_state_search)
  float* out_dist; long long* out_idx;
  int* fail_list;
};

constexpr int SEL_THREADS = 256;
constexpr int RR_THREADS = 256;
constexpr int RR_ROWS = 4;                       // candidate rows in flight per warp
constexpr int RR_BLOCK_ROWS = (RR_THREADS / 32) * RR_ROWS;
constexpr int OUT_THREADS = 128;

// Proven bound on |approximate key - exact key| for one query (DESIGN.md "error bound"), without
// the fp32 rounding of the key itself (3e-7 |key|, added by the callers for the key at hand).
__device__ __forceinline__ double key_error_bound(const FinParams& p, double qn2) {
  const double qn = sqrt(qn2);
  const double bnmax = sqrt((double)__uint_as_float(p.sc->bn2max));
  double eps_dot = (double)p.c_rel * qn * bnmax;
  double eps_cos = (double)p.c_rel * qn;
  if (p.f16) {
    // query elements below fp16's normal range after scaling are quantised to 2^-25
    int sq = f16_scale_exp(__uint_as_float(p.sc->qabsmax)), sbg = 0;
    if (p.direct) direct_exps(p.metric, p.sc, sq, sbg);
    const double qquant = ldexp(1.0, -25 - sq) * sqrt((double)p.d);
    eps_dot += qquant * bnmax;
    eps_cos += qquant;
    if (p.direct) {
      // one scale 2^sbg for the whole base: elements below 2^-14 after scaling are quantised
      // to 2^-25 (scaled units), sum_i |q_i| 2^-25 2^-sbg <= sqrt(d) |q| 2^(-25 - sbg); cosine
      // rows were normalised in fp32 first (two roundings: 2^-23 relative per element)
      const double bquant = ldexp(1.0, -25 - sbg) * sqrt((double)p.d) * qn;
      eps_dot += bquant;
      eps_cos += bquant + ldexp(1.0, -22) * qn;
      // squared L2: |b|^2/2 is accumulated by the tensor core too (same 4-ulp budget per MMA step,
      // one more step) and its three fp16 pieces leave 2^-28 of it
      if (p.metric == NWK_METRIC_SQEUCLIDEAN)
        eps_dot += ((double)p.c_acc + ldexp(1.0, -28)) * 0.5 * bnmax * bnmax + (double)p.c_acc * qn * bnmax;
    }
  }
  if (p.metric == NWK_METRIC_SQEUCLIDEAN) return 2.0 * eps_dot;
  if (p.metric == NWK_METRIC_COSINE) return eps_cos;
  return eps_dot;
}

// Running searches (nwk_topk_state_search): the caller already holds k neighbours of every query
// from earlier slabs of the base; a row of this slab can only enter the result if its exact
// distance is at most the query's current k-th best D.  In key units that is key(D) (the keys omit
// the per-query constants), so rows with an approximate key >= key(D) + eps cannot matter and the
// logging threshold drops to min(tau_q, that): later slabs log and re-rank a fraction of the
// candidates of the first one.  One warp per query.
__global__ void __launch_bounds__(256)
tau_bound_kernel(const FinParams p, const float* __restrict__ bound, long long stride, float* __restrict__ tau,
                 unsigned char* __restrict__ bflag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * 8 + warp;
  if (q >= p.nq) return;
  double qn2 = 0.0;
  const float* qrow = p.query + q * p.d;
  for (int c2 = lane; c2 < p.d; c2 += 32) { const double v = qrow[c2]; qn2 = fma(v, v, qn2); }
  qn2 = warp_sum(qn2);
  if (lane != 0) return;
  unsigned char flag = 0;
  const float D = bound[q * stride];
  if (D < __int_as_float(0x7f800000)) {
    // D is the fp32 rounding of an exact distance: the true value is at most D + ulp(D)/2
    const double d_up = (double)D + fabs((double)D) * 1.2e-7 + 1e-37;
    const double key_shift = (p.metric == NWK_METRIC_SQEUCLIDEAN) ? qn2 : 1.0;
    const double key_scale = (p.metric == NWK_METRIC_COSINE) ? sqrt(qn2) : 1.0;
    const double kb = (d_up - key_shift) * key_scale;
    const double eps = key_error_bound(p, qn2) + 3.0e-7 * fabs(kb) + 1e-30;
    const float t = __double2float_ru(kb + 1.001 * eps);
    if (t < tau[q]) { tau[q] = t; flag = 1; }
  }
  bflag[q] = flag;
}

// select: find the k-th smallest logged key a_k by radix select (4 passes of 8 bits over the
// order-preserving 32-bit keys), form the band [a_k, a_k + 2 eps], certify, and compact the R
// records inside the band to the front of the query's log (order is irrelevant: output_kernel
// sorts by exact distance).  One WARP per query, no block barriers: the log (8 KB on average) is
// re-read from L1 / L2 in every pass, the histogram of a pass is 1 KB of shared memory per warp.
// (The first version used one CTA per query with the log staged in shared memory: 0.6 ms for
// 10 000 queries, half of it waiting at __syncthreads.)
__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const FinParams p) {
  __shared__ unsigned int s_hist_all[SEL_THREADS / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * (SEL_THREADS / 32) + warp;
  if (q >= p.nq) return;                         // whole warps leave; nothing below is block-wide
  unsigned int* const hist = s_hist_all[warp];
  const unsigned int total = p.cnt[q];
  const int c = (int)min(total, (unsigned int)LOG_CAP);
  unsigned long long* const glog = p.log + q * LOG_CAP;
  const float PINF = __int_as_float(0x7f800000);

  // |q|^2 in the same operation order as scan_dist_kernel
  double qn2 = 0.0;
  {
    const float* qrow = p.query + q * p.d;
    for (int c2 = lane; c2 < p.d; c2 += 32) { const double v = qrow[c2]; qn2 = fma(v, v, qn2); }
    qn2 = warp_sum(qn2);
  }
  const bool countable = !(total > (unsigned int)LOG_CAP || c < p.k);

  unsigned int prefix = 0;
  if (countable) {
    // radix select of the element of rank k-1 among the c keys
    unsigned int kk = (unsigned int)(p.k - 1), mask = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0;
      __syncwarp();
      for (int i = lane; i < c; i += 32) {
        unsigned long long rec = glog[i];
        if (shift == 24 && p.direct && !(p.alpha[rec & 0xffffffffull] < PINF)) {
          // direct keys: rows flagged as unrankable (alpha = +inf) were logged with the key of a zero row
          rec |= 0xffffffff00000000ull;
          glog[i] = rec;
        }
        const unsigned int key = (unsigned int)(rec >> 32);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncwarp();
      // lane l owns bins [8l, 8l+8): inclusive scan across lanes, then locate the digit
      unsigned int loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = hist[lane * 8 + j]; sum += loc[j]; }
      unsigned int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned int excl = incl - sum;
      const bool mine = kk >= excl && kk < incl;       // exactly one lane
      unsigned int np = prefix, nk = kk;
      if (mine) {
        unsigned int run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (kk >= run && kk < run + loc[j]) { np = prefix | ((unsigned int)(lane * 8 + j) << shift); nk = kk - run; }
          run += loc[j];
        }
      }
      const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
      prefix = __shfl_sync(0xffffffffu, np, src);
      kk = __shfl_sync(0xffffffffu, nk, src);
      mask |= 255u << shift;
      __syncwarp();
    }
  }

  // every lane evaluates the (warp-uniform) bound: cheaper than a broadcast of its five results
  // A bounded query (running search, tau_q lowered to the caller's k-th distance): every row that
  // can still enter the caller's result is in the log whatever the band looks like, and fewer than
  // k rows may have passed at all -- then all of them are kept and the output is padded.
  const bool bounded = p.bflag != nullptr && p.bflag[q] != 0;
  const bool keep_all = bounded && total <= (unsigned int)LOG_CAP && c < p.k;
  int fail = 0;
  double eps = key_error_bound(p, qn2) + 1e-30;
  unsigned int thr_ord = 0, thr1_ord = 0;
  if (keep_all) {
    thr_ord = thr1_ord = 0xff800000u;        // f2ord(+inf): every logged key (none is nan-ordered unless flagged)
    eps += 3.0e-7 * fabs((double)p.tau[q]);
  } else if (!countable) {
    fail = 1;
  } else {
    const float key_k = ord2f(prefix);       // the k-th smallest approximate key
    // + the fp32 rounding of alpha + beta*acc itself
    eps += 3.0e-7 * fabs((double)key_k);
    const float thr = __double2float_ru((double)key_k + 2.0 * eps);
    const float tau = p.tau[q];
    if (!(thr < tau) && !bounded) fail = 1;  // band reaches the logging threshold: cannot certify
    thr_ord = f2ord(thr);
    thr1_ord = f2ord(__double2float_ru((double)key_k + eps));
  }

  // Compact the records of the inner band (key <= a_k + eps) to the front of the query's log:
  // rerank_kernel recomputes exactly these.  Their exact k-th distance D_k bounds the true k-th
  // from above, so a row of the outer band (a_k + eps, a_k + 2 eps] can only matter when its
  // key is <= D_k + eps; those records are parked at the very end of the log (slots LOG_CAP-1
  // downwards; rerank_kernel's fp64 distances occupy [LOG_CAP/2, LOG_CAP/2 + R)) and
  // output_kernel re-ranks the few that pass that test.  Two sweeps: count, then move in place
  // (a record only ever moves to a slot the warp has already read).
  int R = 0, R2 = 0;
  if (countable || keep_all) {
    for (int i0 = 0; i0 < c; i0 += 32) {
      const int i = i0 + lane;
      const unsigned int o = i < c ? (unsigned int)(glog[i] >> 32) : 0xffffffffu;
      R += __popc(__ballot_sync(0xffffffffu, i < c && o <= thr1_ord));
      R2 += __popc(__ballot_sync(0xffffffffu, i < c && o > thr1_ord && o <= thr_ord));
    }
    // the two regions (and the distances) must not overlap, nor the outer band the unread records
    if (R + R2 > RERANK_MAX || c + R2 > LOG_CAP) fail = 1;
    if (!fail) {
      int w1 = 0, w2 = 0;
      const unsigned int lt = (1u << lane) - 1u;
      for (int i0 = 0; i0 < c; i0 += 32) {
        const int i = i0 + lane;
        const unsigned long long rec = i < c ? glog[i] : 0xffffffffffffffffull;
        const unsigned int o = (unsigned int)(rec >> 32);
        const bool in1 = i < c && o <= thr1_ord;
        const bool in2 = i < c && !in1 && o <= thr_ord;
        const unsigned int b1 = __ballot_sync(0xffffffffu, in1), b2 = __ballot_sync(0xffffffffu, in2);
        if (in1) glog[w1 + __popc(b1 & lt)] = rec;
        if (in2) glog[LOG_CAP - 1 - (w2 + __popc(b2 & lt))] = rec;
        w1 += __popc(b1);
        w2 += __popc(b2);
        __syncwarp();
      }
    }
  }
  if (lane == 0) {
    p.epsq[q] = (float)eps;
    p.qn2[q] = qn2;
    atomicAdd(p.sc->stat_sum + 0, (unsigned long long)total);
    atomicAdd(p.sc->stat_sum + 1, (unsigned long long)min(R, RERANK_MAX));
    atomicMax(p.sc->stat_max + 0, total);
    atomicMax(p.sc->stat_max + 1, (unsigned int)min(R, RERANK_MAX));
    if (fail) {
      int slot = atomicAdd(&p.sc->fail_cnt, 1);
      p.fail_list[slot] = (int)q;
    }
    p.rr[q] = fail ? -1 : R;
    p.rr2[q] = fail ? 0 : R2;
  }
}

// rerank: fp64 distance of every row of the inner band.  HBM gather bound (4*d bytes per row):
// a warp accumulates RR_ROWS rows at a time (they share the query loads and conversions) and,
// before it starts on a batch, asks L2 to fetch the rows of its next batch (prefetch.global.L2,
// no registers or shared memory tied up), so most row reads are L2 hits and the DRAM gather
// runs a batch ahead of the arithmetic.  (Measured alternatives: staging the rows with
// cp.async.bulk, with or without a multi-stage ring, was 10-150 % slower than plain loads.)
__global__ void __launch_bounds__(RR_THREADS, 4)
rerank_kernel(const FinParams p) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  float* s_q = reinterpret_cast<float*>(rr_smem);                                           // [d]
  const long long q = blockIdx.x;
  const int R = p.rr[q];
  const int Y = gridDim.y, y = blockIdx.y;
  if (R <= 0 || y * RR_BLOCK_ROWS >= R) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long* glog = p.log + q * LOG_CAP;
  double* gdist = reinterpret_cast<double*>(p.log + q * LOG_CAP + LOG_CAP / 2);
  const int d = p.d;
  // batch jj of this warp covers the log entries [first(jj), first(jj) + RR_ROWS)
  auto first = [&](int jj) { return ((jj * Y + y) * (RR_THREADS / 32) + warp) * RR_ROWS; };
  auto load_id = [&](int jj) -> int {                        // row ids fit 31 bits (tc_path_supported)
    const int i = first(jj) + lane;
    return (lane < RR_ROWS && i < R) ? (int)(glog[i] & 0xffffffffull) : -1;
  };
  auto prefetch = [&](int ids) {
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      const int row = __shfl_sync(0xffffffffu, ids, r);
      if (row >= 0)
        for (int off = lane * 32; off < d; off += 32 * 32)     // one 128-byte line per lane
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.base + (long long)row * d + off));
    }
  };
  int id_cur = load_id(0);
  int id_nxt = load_id(1);
  prefetch(id_cur);
  for (int i = threadIdx.x * 4; i < d; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(s_q + i) = __ldg(reinterpret_cast<const float4*>(p.query + q * d + i));
  __syncthreads();
  const double qn2 = p.qn2[q];

  for (int jj = 0; first(jj) < R; ++jj) {                    // warp-uniform
    prefetch(id_nxt);
    const int id_nn = load_id(jj + 2);
    const int i0 = first(jj);
    int rows[RR_ROWS];
    double acc[RR_ROWS], bn2[RR_ROWS];
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      acc[r] = 0.0; bn2[r] = 0.0;
      rows[r] = __shfl_sync(0xffffffffu, id_cur, r);
    }
    for (int c2 = lane * 4; c2 < d; c2 += 128) {
      const float4 q4 = *reinterpret_cast<const float4*>(s_q + c2);
      float4 b4[RR_ROWS];
#pragma unroll
      for (int r = 0; r < RR_ROWS; ++r)
        b4[r] = rows[r] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.base + (long long)rows[r] * d + c2))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < RR_ROWS; ++r) {
        // same operation order as scan_dist_kernel: both exact paths agree bit for bit
        // (the conversions stay F2F here: moving them to the integer pipe with f2d_exact, which pays in the
        //  DFMA-bound MaxSim re-scoring, made this HBM-gather-bound kernel slower -- finish 17.8 -> 20.8 ms on a
        //  config-4 slab, 1.49 -> 1.59 ms at config 2)
        const double b0 = b4[r].x, b1 = b4[r].y, b2 = b4[r].z, b3 = b4[r].w;
        if (p.metric == NWK_METRIC_COSINE) {
          bn2[r] = fma(b0, b0, bn2[r]); bn2[r] = fma(b1, b1, bn2[r]);
          bn2[r] = fma(b2, b2, bn2[r]); bn2[r] = fma(b3, b3, bn2[r]);
        }
        if (p.metric == NWK_METRIC_SQEUCLIDEAN) {
          const double e0 = (double)q4.x - b0, e1 = (double)q4.y - b1, e2 = (double)q4.z - b2, e3 = (double)q4.w - b3;
          acc[r] = fma(e0, e0, acc[r]); acc[r] = fma(e1, e1, acc[r]);
          acc[r] = fma(e2, e2, acc[r]); acc[r] = fma(e3, e3, acc[r]);
        } else {
          acc[r] = fma((double)q4.x, b0, acc[r]); acc[r] = fma((double)q4.y, b1, acc[r]);
          acc[r] = fma((double)q4.z, b2, acc[r]); acc[r] = fma((double)q4.w, b3, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      double dist = warp_sum(acc[r]);
      if (p.metric == NWK_METRIC_COSINE) bn2[r] = warp_sum(bn2[r]);
      if (p.metric != NWK_METRIC_SQEUCLIDEAN) dist = finish_metric(p.metric, dist, qn2, bn2[r]);
      if (lane == 0 && rows[r] >= 0) gdist[i0 + r] = dist;
    }
    id_cur = id_nxt;
    id_nxt = id_nn;
  }
}

// fp64 distance of one base row to the query by one warp; the operation order (float4 per lane,
// stride 128, one fma chain per lane, xor-shuffle tree) is the one every exact path uses
__device__ __forceinline__ double warp_row_distance(const float* __restrict__ qv, const float* __restrict__ bv,
                                                    int d, int metric, double qn2, int lane) {
  double acc = 0.0, bn2 = 0.0;
  for (int c2 = lane * 4; c2 < d; c2 += 128) {
    const float4 q4 = *reinterpret_cast<const float4*>(qv + c2);
    const float4 b4 = *reinterpret_cast<const float4*>(bv + c2);
    const double b0 = b4.x, b1 = b4.y, b2 = b4.z, b3 = b4.w;
    if (metric == NWK_METRIC_COSINE) {
      bn2 = fma(b0, b0, bn2); bn2 = fma(b1, b1, bn2);
      bn2 = fma(b2, b2, bn2); bn2 = fma(b3, b3, bn2);
    }
    if (metric == NWK_METRIC_SQEUCLIDEAN) {
      const double e0 = (double)q4.x - b0, e1 = (double)q4.y - b1, e2 = (double)q4.z - b2, e3 = (double)q4.w - b3;
      acc = fma(e0, e0, acc); acc = fma(e1, e1, acc);
      acc = fma(e2, e2, acc); acc = fma(e3, e3, acc);
    } else {
      acc = fma((double)q4.x, b0, acc); acc = fma((double)q4.y, b1, acc);
      acc = fma((double)q4.z, b2, acc); acc = fma((double)q4.w, b3, acc);
    }
  }
  double dist = warp_sum(acc);
  if (metric == NWK_METRIC_COSINE) bn2 = warp_sum(bn2);
  if (metric != NWK_METRIC_SQEUCLIDEAN) dist = finish_metric(metric, dist, qn2, bn2);
  return dist;
}

// output: order the re-ranked rows by (fp64 distance, index), pull in the few rows of the outer
// band that can still matter, emit k, monitor the error bound
__global__ void __launch_bounds__(OUT_THREADS)
output_kernel(const FinParams p) {
  __shared__ double s_key[RERANK_MAX];
  __shared__ long long s_id[RERANK_MAX];
  __shared__ float s_akey[64];                 // approximate keys of the outer-band rows taken in
  __shared__ float s_ratio[OUT_THREADS / 32];
  __shared__ int s_nx;
  const long long q = blockIdx.x;
  const int R = p.rr[q];
  if (R < 0) return;                           // uncertified: the exact scan writes this query's row
  if (R == 0) {
    // a bounded query none of whose rows in this slab can enter the caller's result
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
      p.out_dist[q * p.k + j] = __int_as_float(0x7f800000);
      p.out_idx[q * p.k + j] = -1;
    }
    return;
  }
  const int R2 = p.rr2[q];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int r2 = max(next_pow2(max(R, p.k)), 2);     // (a bounded query may hold fewer than k rows: padded with +inf / -1)
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const unsigned long long* glog = p.log + q * LOG_CAP;
  const double* gdist = reinterpret_cast<const double*>(p.log + q * LOG_CAP + LOG_CAP / 2);
  const double qn2 = p.qn2[q];
  const double eps = (double)p.epsq[q];
  // exact key = (distance - key_shift) * key_scale: the candidate keys omit the per-query
  // constants (|q|^2 for sq-L2, the 1 of cosine / dot, and cosine's 1/|q|)
  const double key_shift = (p.metric == NWK_METRIC_SQEUCLIDEAN) ? qn2 : 1.0;
  const double key_scale = (p.metric == NWK_METRIC_COSINE) ? sqrt(qn2) : 1.0;
  float ratio = 0.f;
  if (threadIdx.x == 0) s_nx = 0;
  for (int i = threadIdx.x; i < r2; i += blockDim.x) {
    double dist = INF;
    long long id = 0x7fffffffffffffffLL;
    if (i < R) {
      const unsigned long long rec = glog[i];
      id = (long long)(rec & 0xffffffffull);
      dist = gdist[i];
      if (dist == dist) {
        const float akey = ord2f((uint32_t)(rec >> 32));
        ratio = fmaxf(ratio, (float)(fabs((double)akey - (dist - key_shift) * key_scale) / eps));
      } else {
        dist = INF;
      }
    }
    s_key[i] = dist;
    s_id[i] = id;
  }
  block_bitonic_sort(s_key, s_id, r2);

  if (R2 > 0) {
    // D_k = exact k-th distance among the inner band >= the true k-th distance, so a row outside
    // the inner band belongs to the answer only if its approximate key is <= key(D_k) + eps
    const double bound = (s_key[p.k - 1] - key_shift) * key_scale + eps;
    __syncthreads();                           // every thread has read s_key[k-1] before the tail is reused
    for (int i = threadIdx.x; i < R2; i += blockDim.x) {
      const unsigned long long rec = glog[LOG_CAP - 1 - i];
      const float akey = ord2f((uint32_t)(rec >> 32));
      if ((double)akey <= bound) {
        const int pos = atomicAdd(&s_nx, 1);
        if (pos < 64 && R + pos < RERANK_MAX) {
          s_id[R + pos] = (long long)(rec & 0xffffffffull);
          s_akey[pos] = akey;
        }
      }
    }
    __syncthreads();
    const int nx = s_nx;
    if (nx > 0) {
      if (nx > 64 || R + nx > RERANK_MAX) {
        // cannot happen unless the inner band's k-th distance sits a whole eps above a_k for
        // dozens of rows: hand the query to the exact scan
        if (threadIdx.x == 0) {
          const int slot = atomicAdd(&p.sc->fail_cnt, 1);
          p.fail_list[slot] = (int)q;
        }
        return;
      }
      for (int e = warp; e < nx; e += OUT_THREADS / 32) {
        const long long row = s_id[R + e];
        double dist = warp_row_distance(p.query + q * p.d, p.base + row * p.d, p.d, p.metric, qn2, lane);
        if (lane == 0) {
          if (dist == dist) ratio = fmaxf(ratio, (float)(fabs((double)s_akey[e] - (dist - key_shift) * key_scale) / eps));
          else dist = INF;
          s_key[R + e] = dist;
        }
      }
      const int r3 = max(next_pow2(R + nx), 2);
      __syncthreads();
      for (int i = R + nx + threadIdx.x; i < r3; i += blockDim.x) { s_key[i] = INF; s_id[i] = 0x7fffffffffffffffLL; }
      r2 = r3;
      block_bitonic_sort(s_key, s_id, r2);
      if (threadIdx.x == 0) atomicAdd(p.sc->stat_sum + 1, (unsigned long long)nx);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ratio = fmaxf(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
  if (lane == 0) s_ratio[warp] = ratio;
  __syncthreads();
  for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
    const bool pad = s_id[j] == 0x7fffffffffffffffLL;
    p.out_dist[q * p.k + j] = (float)s_key[j];
    p.out_idx[q * p.k + j] = pad ? -1 : s_id[j] + p.id_offset;
  }
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < OUT_THREADS / 32; ++w) r = fmaxf(r, s_ratio[w]);
    atomicMax(&p.sc->err_ratio, __float_as_uint(r));
  }
}

// ---------------------------------------------------------------------------------------
// second chance: the queries of a fail list as a compact query set, and their results back
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
retry_gather_kernel(const float* __restrict__ query, const int* __restrict__ list, int n, int d,
                    const float* __restrict__ bound, long long bound_stride, float* __restrict__ rq,
                    float* __restrict__ rbound, int* __restrict__ rlist) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + warp; i < n; i += gridDim.x * 8) {
    const long long q = list[i];
    for (int c = lane; c < d; c += 32) rq[(long long)i * d + c] = query[q * d + c];
    if (lane == 0) {
      rlist[i] = (int)q;
      if (bound) rbound[i] = bound[q * bound_stride];
    }
  }
}

__global__ void __launch_bounds__(256)
retry_scatter_kernel(const float* __restrict__ rd, const long long* __restrict__ ri, const int* __restrict__ rlist, int n,
                     int k, float* __restrict__ out_dist, long long* __restrict__ out_idx) {
  const long long total = (long long)n * k;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / k), j = (int)(t - (long long)i * k);
    const long long q = rlist[i];
    out_dist[q * k + j] = rd[t];
    out_idx[q * k + j] = ri[t];
  }
}

// a retry keeps the scales and the prepared base of its parent call: only the per-call counters start over
__global__ void retry_reset_kernel(Scalars* sc) {
  sc->fail_cnt = 0;
  sc->stat_max[0] = sc->stat_max[1] = 0u;
  sc->stat_sum[0] = sc->stat_sum[1] = 0ull;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym)
      return fail(NWK_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  *out = fn;
  return NWK_OK;
}

// 2-D row-major operand [rows, row_elems] (elements of 4 or 2 bytes); box = 128 bytes x box_rows
static int make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long row_elems, bool f16,
                     int box_rows, long long pitch_elems = 0) {
  EncodeTiledFn enc;
  NWK_TRY(get_encode_fn(&enc));
  const int esz = f16 ? 2 : 4;
  if (pitch_elems <= 0) pitch_elems = row_elems;     // columns in [row_elems, pitch) are never read: TMA zero-fills them
  cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(NWK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NWK_OK;
}

static int tc_optin() {
  DevState* ds = dev_state();          // cudaFuncSetAttribute is per device
  if (!ds->tc_optin) {
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_SAMPLE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_DUMP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_SAMPLE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_FILTER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    NWK_CUDA(cudaFuncSetAttribute(tc_kernel<MODE_DUMP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    ds->tc_optin = 1;
  }
  return NWK_OK;
}

bool tc_path_supported(long long nb, long long nq, int d, int k) {
  return (d % 4 == 0) && d >= 4 && k <= TC_MAX_K && nb >= 1 && nq >= 1 &&
         nb < (1ll << 31) && nq < (1ll << 31) && (long long)d * 4 < (1ll << 31);
}

struct TcPlan {
  int nQT, nT, nKB, kLast, G, nChunks, nUnits, nS, sChunks, sParts, dh, kcols, pitch, sM;
  bool retry_ok;            // the first pass uses a rank below SAMPLE_M: uncertified queries get a second chance at SAMPLE_M
  bool sample;
  long long nb_pad, nq_pad;
  long long s_stride, ns, ns_pad;
};

// xcols = extra K columns of the fp16 operand copies (L2X_COLS for direct squared-L2 keys)
static TcPlan make_plan(long long nb, long long nq, int d, int k, int sms, bool f16, int xcols = 0, int first_rank_override = 0) {
  TcPlan pl{};
  const int bk = f16 ? 64 : 32;
  pl.dh = (d + 7) / 8 * 8;
  // rows of the fp16 copies start on 128-byte lines (a box row that straddles two lines doubles the
  // L2 -> SM requests: measured +50 % filter time at d = 1536 with a pitch of 1544)
  pl.kcols = f16 ? pl.dh + xcols : d;                // row elements the MMAs have to cover
  pl.pitch = f16 ? (pl.kcols + 63) / 64 * 64 : d;
  pl.nQT = (int)((nq + TQ - 1) / TQ);
  pl.nT = (int)((nb + TB - 1) / TB);
  const int kcols = pl.kcols;
  pl.nKB = (kcols + bk - 1) / bk;
  pl.kLast = (kcols - (pl.nKB - 1) * bk + (f16 ? 15 : 7)) / (f16 ? 16 : 8);
  pl.nb_pad = (long long)pl.nT * TB;
  pl.nq_pad = (long long)pl.nQT * TQ;
  long long pairs = (long long)pl.nQT * pl.nT;
  long long g = pairs / ((long long)sms * 8);
  pl.G = (int)(g < 1 ? 1 : (g > 16 ? 16 : g));
  pl.nChunks = (pl.nT + pl.G - 1) / pl.G;
  pl.nUnits = pl.nChunks * pl.nQT;
  // sampling only pays when the base is larger than what the log can simply hold
  pl.sample = nb > LOG_CAP / 2;
  pl.sM = SAMPLE_M;
  pl.retry_ok = false;
  if (pl.sample) {
    // the SAMPLE_M-th smallest key of a sample of every `stride`-th row has about
    // SAMPLE_M * stride rows of the whole base below it
    // The threshold statistic: the sM-th smallest key of a stratified sample of every s_stride-th row has
    // about sM * s_stride rows of the whole base below it, with a relative spread of 1 / sqrt(sM).
    //   long rows  (d >= 512): every 64th row, rank 16 -> ~1024 candidates per query; the sample pass costs
    //              1/64 of the filter pass's FLOPs (measured -0.3 ms of 13.3 at config 2 against 32 / 32);
    //   short rows: every 32nd row, rank 12 -> ~384 candidates: there the filter's epilogue is the limit
    //              and every candidate costs a trip through its slow path (and 8 bytes of log traffic).
    // The rank grows with k so that about 3.5 k rows pass.  A low rank leaves the odd query with a threshold
    // below what its band needs (2 of 10 000 at rank 16 on a 5M x 1536 shard); such a query used to cost an
    // fp64 scan of the whole base (+32 ms there) and now gets a SECOND CHANCE: the same sample at rank
    // SAMPLE_M (twice to three times the candidates), one more tensor-core pass over the base for the
    // handful of uncertified queries (tc_knn's retry).  Only what fails again goes to the scan.
    // Small bases keep the dense, high-rank statistic: their sample pass costs microseconds, while a second
    // chance costs a dozen launches and a host round trip (measured at config 1, 1000 x 10 000 x 384:
    // 0.33 ms per call with rank 32, 0.43 ms with rank 12 and its 4 second-chance queries).
    const bool small = nb < 200000;
    pl.s_stride = (d >= 512 && !small) ? 64 : 32;
    int rank = small ? SAMPLE_M : (d >= 512 ? 16 : 12);
    const int need = (int)((7ll * k / 2 + pl.s_stride - 1) / pl.s_stride);
    if (rank < need) rank = need;
    if (const char* e = getenv("NWK_SAMPLE_M")) rank = atoi(e);                              // experiment switches
    if (const char* e = getenv("NWK_SAMPLE_STRIDE")) pl.s_stride = atoi(e) >= 64 ? 64 : 32;
    rank = (rank + 3) / 4 * 4;
    pl.sM = rank < 8 ? 8 : (rank > SAMPLE_M ? SAMPLE_M : rank);
    if (first_rank_override > 0) pl.sM = first_rank_override;
    pl.retry_ok = pl.sM < SAMPLE_M;
    pl.ns = (nb + pl.s_stride - 1) / pl.s_stride;
    pl.ns_pad = (pl.ns + TB - 1) / TB * TB;
    pl.nS = (int)(pl.ns_pad / TB);
    pl.sChunks = TB / 32;
    // split each query tile's scan of the sample into parts so that all SMs work: pick the
    // part count with the fewest tile-times on the critical path (every part pays ~3 tile-times
    // of list warm-up and publication)
    pl.sParts = 1;
    long long best = -1;
    for (int parts = 1; parts <= SAMPLE_MAX_PARTS && parts <= pl.nS; ++parts) {
      const long long waves = ((long long)pl.nQT * parts + sms - 1) / sms;
      const long long cost = waves * ((pl.nS + parts - 1) / parts + 3);
      if (best < 0 || cost < best) { best = cost; pl.sParts = parts; }
    }
  }
  return pl;
}

// Workspace layout: [base-dependent buffers][retry area][query-dependent buffers][fp64 scan scratch].
// The first two depend on (nb, d, k) only, so the second-chance call (same base, a few queries) finds the
// prepared base where its parent left it and may reuse the parent's query-dependent region.
struct TcArena {
  // base-dependent
  float *alpha, *beta, *salpha, *sbeta, *srows;
  __half* bh;
  Scalars* sc;
  // retry area
  float *rq, *rbound, *rout_d;
  long long* rout_i;
  int* rlist;
  // query-dependent
  float *tau, *epsq, *slist;
  unsigned int* cnt;
  unsigned long long* log;
  int *fail_list, *rr, *rr2;
  double* qn2;
  unsigned char* bflag;
  __half* qh;
  size_t end;
};

static TcArena carve(void* ws, size_t ws_bytes, const TcPlan& pl, long long nb, long long nq, int d, int k) {
  Arena ar(ws, ws_bytes);
  TcArena a{};
  a.alpha = ar.take<float>((size_t)pl.nb_pad);
  a.beta = ar.take<float>((size_t)pl.nb_pad);
  a.bh = ar.take<__half>((size_t)nb * pl.pitch);
  a.salpha = ar.take<float>((size_t)pl.ns_pad);
  a.sbeta = ar.take<float>((size_t)pl.ns_pad);
  a.srows = ar.take<float>((size_t)(pl.ns + 1) * (size_t)(d > pl.pitch ? d : pl.pitch));
  a.sc = ar.take<Scalars>(1);
  a.rq = ar.take<float>((size_t)RETRY_MAX * d);
  a.rbound = ar.take<float>((size_t)RETRY_MAX);
  a.rout_d = ar.take<float>((size_t)RETRY_MAX * k);
  a.rout_i = ar.take<long long>((size_t)RETRY_MAX * k);
  a.rlist = ar.take<int>((size_t)RETRY_MAX);
  a.tau = ar.take<float>((size_t)pl.nq_pad);
  a.cnt = ar.take<unsigned int>((size_t)nq);
  a.log = ar.take<unsigned long long>((size_t)nq * LOG_CAP);
  a.fail_list = ar.take<int>((size_t)nq);
  a.rr = ar.take<int>((size_t)nq);
  a.rr2 = ar.take<int>((size_t)nq);
  a.epsq = ar.take<float>((size_t)nq);
  a.qn2 = ar.take<double>((size_t)nq);
  a.bflag = ar.take<unsigned char>((size_t)nq);
  a.slist = ar.take<float>((size_t)nq * SAMPLE_MAX_LISTS * SAMPLE_M);
  a.qh = ar.take<__half>((size_t)nq * pl.pitch);
  a.end = align_up(ar.off, 256);
  return a;
}

size_t tc_workspace_bytes(long long nb, long long nq, int d, int k) {
  // (the widest variant: direct squared-L2 keys add L2X_COLS columns to the operand copies)
  TcPlan pl = make_plan(nb, nq, d, k, 148, true, L2X_COLS);
  const TcArena a = carve(nullptr, 0, pl, nb, nq, d, k);
  // fallback scan for up to 64 queries is always provisioned; more is handled in rounds
  return a.end + exact_scan_workspace_bytes(nb, 64, k) + 256 + 4096;
}

template <bool F16>
static int launch_tc(int mode, int grid, const CUtensorMap& tq, const CUtensorMap& tb, const TcParams& p,
                     cudaStream_t s) {
  if (mode == MODE_SAMPLE) tc_kernel<MODE_SAMPLE, F16><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tq, tb, p);
  else if (mode == MODE_FILTER) tc_kernel<MODE_FILTER, F16><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tq, tb, p);
  else tc_kernel<MODE_DUMP, F16><<<grid, 128 + 128 * epi_wg<MODE_DUMP, false, false>(), TC_SMEM_BYTES, s>>>(tq, tb, p);
  NWK_LAUNCH_CHECK();
  return NWK_OK;
}

int tc_knn(const float* base, long long nb, const float* query, long long nq, int d, int k, int metric,
           long long id_offset, float* out_dist, long long* out_idx, void* ws, size_t ws_bytes,
           unsigned int flags, nwk_knn_stats* stats, cudaStream_t s, const float* bound, long long bound_stride) {
  NWK_TRY(tc_optin());
  std::lock_guard<std::recursive_mutex> dev_lock(dev_state()->mu);
  const int sms = sm_count();
  const bool f16 = !(flags & NWK_FLAG_TF32);
  // Direct keys (no alpha / beta in the epilogue).  Cosine rows are normalised in the preparation
  // pass anyway, so its direct keys cost nothing extra and pay at any d (-6 % filter time at
  // d = 768 under the power cap).  Dot needs one more read of the base for the global scale, and
  // squared L2 carries |b|^2 in L2X_COLS extra K columns = one more K-block per tile (with a scale
  // taken from the queries; NWK_FLAG_ALPHA_BETA is the repeat after a scale overflow): both only
  // pay for short rows, where the epilogue is the limit (measured for squared L2: -7.5 % per step
  // at d = 128, +1.5 % at d = 1536).
  bool direct = f16 && !(flags & NWK_FLAG_ALPHA_BETA) && (metric == NWK_METRIC_COSINE || (d + 7) / 8 * 8 <= 256);
  if (const char* e = getenv("NWK_DIRECT")) direct = f16 && !(flags & NWK_FLAG_ALPHA_BETA) && atoi(e) != 0;   // experiment switch
  const bool l2x = direct && metric == NWK_METRIC_SQEUCLIDEAN;
  const bool retry = (flags & FLAG_INTERNAL_RETRY) != 0;
  TcPlan pl = make_plan(nb, nq, d, k, sms, f16, l2x ? L2X_COLS : 0, retry ? SAMPLE_M : 0);
  const TcArena wa = carve(ws, ws_bytes, pl, nb, nq, d, k);
  float *alpha = wa.alpha, *beta = wa.beta, *tau = wa.tau, *epsq = wa.epsq, *slist = wa.slist;
  float *salpha = wa.salpha, *sbeta = wa.sbeta, *srows = wa.srows;
  unsigned int* cnt = wa.cnt;
  unsigned long long* log = wa.log;
  int *fail_list = wa.fail_list, *rr = wa.rr, *rr2 = wa.rr2;
  double* qn2 = wa.qn2;
  unsigned char* bflag = wa.bflag;
  __half *bh = wa.bh, *qh = wa.qh;
  Scalars* sc = wa.sc;
  if (wa.end > ws_bytes) return fail(NWK_ERR_WORKSPACE, "workspace too small for the tensor-core path");
  void* scan_ws = static_cast<char*>(ws) + wa.end;
  size_t scan_ws_bytes = ws_bytes - wa.end;

  // Filter variant.  CTA pairs (cta_group::2) when the query tiles pair up without much waste and
  // the MMAs of a tile outlast its epilogue (measured: -6 % filter time at d = 768, -10 % at
  // d = 1536, but +25 % at d = 128, where the epilogue is the limit and the pair couples two of
  // them).
  bool pair = f16 && sms >= 2 && d >= 512 && (pl.nQT % 2 == 0 || pl.nQT >= 8);
  if (const char* e = getenv("NWK_PAIR")) pair = f16 && sms >= 2 && atoi(e) != 0;   // experiment switch
  // resident query tiles: single-CTA filter, rows of at most 6 K-blocks (d <= 384)
  bool ares = f16 && !pair && pl.nKB <= 6;
  if (const char* e = getenv("NWK_ARES")) ares = f16 && !pair && pl.nKB <= 6 && atoi(e) != 0;

  const long long launches0 = g_launches.load();
  const bool profile = stats && (flags & NWK_FLAG_PROFILE);
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (profile) {
    for (int i = 0; i < 6; ++i) cudaEventCreate(&ev[i]);
    cudaEventRecord(ev[0], s);
  }
  if (retry) {
    // second chance: the scales in `sc`, the fp16 base, alpha / beta and the sample matrix are the parent call's
    retry_reset_kernel<<<1, 1, 0, s>>>(sc);
    NWK_LAUNCH_CHECK();
  } else {
    NWK_CUDA(cudaMemsetAsync(sc, 0, sizeof(Scalars), s));
  }
  NWK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nq * 4, s));

  if (f16) {
    if (!retry) {
      qmax_kernel<<<sms, 256, 0, s>>>(query, nq * d, &sc->qabsmax);
      NWK_LAUNCH_CHECK();
      if (l2x) {
        qnorm_kernel<<<sms * 2, 256, 0, s>>>(query, nq, d, &sc->qn2max);
        NWK_LAUNCH_CHECK();
      }
    }
    qprep_kernel<<<sms * 2, 256, 0, s>>>(query, nq, d, pl.dh, pl.pitch, metric, direct ? 1 : 0, sc, qh);
    NWK_LAUNCH_CHECK();
  }
  if (direct && metric == NWK_METRIC_DOT && !retry) {
    // the one extra pass the direct keys cost: the global scale needs max |b_ij| before any row is converted
    qmax_kernel<<<sms * 8, 256, 0, s>>>(base, nb * d, &sc->babsmax);
    NWK_LAUNCH_CHECK();
  }
  SampleOut no_sample{};
  no_sample.stride_log2 = -1;
  if (pl.sample && !retry) {
    // the sample rows first (1/32 .. 1/64 of the base): the sample pass can start while the bulk of the
    // base is still being prepared
    SampleOut so{};
    so.stride_log2 = 0;
    while ((1ll << so.stride_log2) < pl.s_stride) ++so.stride_log2;     // a power of two (make_plan)
    so.ns = pl.ns; so.ns_pad = pl.ns_pad;
    so.rows = srows; so.alpha = salpha; so.beta = sbeta;
    long long blocks = std::min<long long>((pl.ns_pad + 7) / 8, (long long)sms * 16);
    if (f16 && d <= 256) baseprep_kernel<true, 2><<<(unsigned)blocks, 256, 0, s>>>(base, nb, 0, 0, d, pl.dh, metric, alpha, beta, sc, bh, so, direct ? 1 : 0, pl.pitch);
    else if (f16) baseprep_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(base, nb, 0, 0, d, pl.dh, metric, alpha, beta, sc, bh, so, direct ? 1 : 0, pl.pitch);
    else baseprep_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(base, nb, 0, 0, d, pl.dh, metric, alpha, beta, sc, bh, so, 0, pl.dh);
    NWK_LAUNCH_CHECK();
  }

  // ---- the base in parts: preparation (HBM-bound) on a side stream, filter launches (tensor-bound) behind it.
  // Part p's filter pass waits only for part p's preparation, so after the first part everything the
  // preparation costs is hidden under the filter pass; baseprep CTAs use no shared memory and few
  // registers, they run beside the persistent filter CTA of an SM.  A caller's distance bound
  // (running searches) needs max |b|^2 of the whole slab before the thresholds exist: no overlap there
  // beyond the sample pass.
  DevState* ds = dev_state();
  // tiles per part: multiples of the schedule's chunk of G tiles
  const int workers = pair ? sms / 2 : sms;
  const int nQTf = pair ? (pl.nQT + 1) / 2 : pl.nQT;          // query tiles (or tile pairs) of the filter schedule
  int G = pl.G;
  if (pair) {
    long long g = (long long)nQTf * pl.nT / ((long long)workers * 8);
    G = (int)(g < 1 ? 1 : (g > 16 ? 16 : g));
  }
  int nparts = 1;
  {
    // every part must still give each worker several units, or its tail costs more than the overlap wins
    const long long units_total = (long long)((pl.nT + G - 1) / G) * nQTf;
    // (the 640-thread direct-key variants leave no registers for a second CTA on their SM: no parts there)
    const bool roomy = pair || !direct;
    if (roomy && units_total >= (long long)workers * 24) nparts = 3;
    else if (roomy && units_total >= (long long)workers * 12) nparts = 2;
    if (const char* e = getenv("NWK_PARTS")) nparts = std::max(1, std::min(8, atoi(e)));     // experiment switch
    if (nparts > (pl.nT + G - 1) / G || retry) nparts = 1;
  }
  int part_t0[9];
  {
    const int chunks = (pl.nT + G - 1) / G;
    // a short first part (the only preparation that is exposed), the rest evenly
    part_t0[0] = 0;
    for (int pi = 1; pi <= nparts; ++pi) {
      long long c = nparts == 1 ? chunks : (long long)chunks * (2 * pi - 1) / (2 * nparts - 1);
      if (const char* e = getenv("NWK_PART0_PCT"))                                          // experiment switch
        c = (long long)chunks * atoi(e) / 100 + (long long)(chunks - chunks * atoi(e) / 100) * (pi - 1) / (nparts - 1 > 0 ? nparts - 1 : 1);
      if (pi == nparts) c = chunks;
      part_t0[pi] = (int)std::min<long long>((long long)c * G, pl.nT);
    }
  }
  const bool prep_aside = nparts > 1;
  cudaStream_t sp = prep_aside ? ds->side[0] : s;
  if (prep_aside) {
    NWK_CUDA(cudaEventRecord(ds->ev_fork, s));
    NWK_CUDA(cudaStreamWaitEvent(sp, ds->ev_fork, 0));
  }
  auto launch_prep = [&](int pi) -> int {
    const long long r0 = (long long)part_t0[pi] * TB;
    const long long r1 = pi == nparts - 1 ? pl.nb_pad : (long long)part_t0[pi + 1] * TB;
    // beside a 384-thread filter CTA (12 warps x 4096 registers) an SM has 16 K registers left: 4-warp
    // CTAs of this 80-register kernel fit, 8-warp CTAs do not
    const int threads = prep_aside ? 128 : 256;
    long long blocks = std::min<long long>((r1 - r0) / (threads / 32) + 1, (long long)sms * (prep_aside ? 32 : 16));
    if (f16 && d <= 256) baseprep_kernel<true, 2><<<(unsigned)blocks, threads, 0, sp>>>(base, nb, r0, r1, d, pl.dh, metric, alpha, beta, sc, bh, no_sample, direct ? 1 : 0, pl.pitch);
    else if (f16) baseprep_kernel<true><<<(unsigned)blocks, threads, 0, sp>>>(base, nb, r0, r1, d, pl.dh, metric, alpha, beta, sc, bh, no_sample, direct ? 1 : 0, pl.pitch);
    else baseprep_kernel<false><<<(unsigned)blocks, threads, 0, sp>>>(base, nb, r0, r1, d, pl.dh, metric, alpha, beta, sc, bh, no_sample, 0, pl.dh);
    NWK_LAUNCH_CHECK();
    if (prep_aside) NWK_CUDA(cudaEventRecord(ds->ev_chunk[pi], sp));
    return NWK_OK;
  };
  // All parts are enqueued at once: the preparation runs beside the sample pass and the first filter launch.
  // (Holding parts 1.. back until the thresholds exist -- NWK_PREP_GATE=1 -- shortens the sample pass, 0.78 ->
  // 0.49 ms at config 2, and lengthens the filter pass by as much, 10.24 -> 10.62 ms: the HBM traffic of the
  // preparation costs the tensor kernels the same time wherever it is placed.  Measured, profiles/r02_summary.md.)
  bool gate = false;
  if (const char* e = getenv("NWK_PREP_GATE")) gate = prep_aside && pl.sample && !bound && atoi(e) != 0;    // experiment switch
  for (int pi = 0; pi < (gate ? 1 : nparts) && !retry; ++pi) NWK_TRY(launch_prep(pi));
  if (profile) cudaEventRecord(ev[1], s);

  CUtensorMap tm_q, tm_b;
  if (f16) {
    NWK_TRY(make_tmap(&tm_q, qh, nq, pl.kcols, true, TQ, pl.pitch));
    NWK_TRY(make_tmap(&tm_b, bh, nb, pl.kcols, true, TB, pl.pitch));
  } else {
    NWK_TRY(make_tmap(&tm_q, query, nq, d, false, TQ));
    NWK_TRY(make_tmap(&tm_b, base, nb, d, false, TB));
  }

  TcParams p{};
  p.nKB = pl.nKB; p.kLast = pl.kLast; p.nq = nq; p.nb = nb; p.nQT = pl.nQT; p.nT = pl.nT;
  p.G = pl.G; p.nChunks = pl.nChunks; p.nUnits = pl.nUnits; p.nS = pl.nS; p.sChunks = pl.sChunks; p.sParts = pl.sParts;
  p.alpha = alpha; p.beta = beta; p.tau = tau; p.slist = slist; p.cnt = cnt; p.log = log; p.err = &sc->err;
  p.sc = sc; p.metric = metric; p.sM = pl.sM;
  unsigned long long* d_prof = nullptr;
#ifdef NWK_EXPERIMENTS
  // timing experiments (scripts/gpu_exp.sh builds libnwk_exp.so with -DNWK_EXPERIMENTS); the
  // product library does not read these switches: NWK_DEBUG_EPI makes the results wrong
  if (const char* e = getenv("NWK_DEBUG_EPI")) p.dbg = atoi(e);
  if (const char* e = getenv("NWK_DEBUG_NA")) p.dbg_na = atoi(e);
  if (getenv("NWK_DEBUG_PROF")) {
    NWK_CUDA(cudaMalloc(&d_prof, (size_t)sms * 4 * 8));
    NWK_CUDA(cudaMemsetAsync(d_prof, 0, (size_t)sms * 4 * 8, s));
    p.prof = d_prof;
  }
#endif

  if (pl.sample) {
    fill_f32_kernel<<<64, 256, 0, s>>>(slist, nq * pl.sParts * SAMPLE_WG * SAMPLE_M, HUGE_VALF);
    NWK_LAUNCH_CHECK();
    // the sample pass scans the compact matrix of sampled rows with their own alpha / beta
    CUtensorMap tm_s;
    if (f16) NWK_TRY(make_tmap(&tm_s, srows, pl.ns, pl.kcols, true, TB, pl.pitch));
    else NWK_TRY(make_tmap(&tm_s, srows, pl.ns, d, false, TB));
    TcParams ps = p;
    ps.alpha = salpha; ps.beta = sbeta; ps.nT = pl.nS;
    // The warm-up threshold (4th smallest key of a part's first 128 sample rows) must stay above
    // the final tau (the sM-th smallest of all ns rows), or list entries go missing and tau
    // comes out too large: the chance per list is P(Poisson(128 * sM / ns) >= 4), 0.2 % at
    // ns = 8192 and 1e-5 at the 31 250 rows of a 1M-row base (sM = 32); smaller samples insert every key.
    ps.sWarm = pl.ns >= 8192 ? 1 : 0;
    int grid = pl.sParts * pl.nQT < sms ? pl.sParts * pl.nQT : sms;
    NWK_TRY(f16 ? launch_tc<true>(MODE_SAMPLE, grid, tm_q, tm_s, ps, s) : launch_tc<false>(MODE_SAMPLE, grid, tm_q, tm_s, ps, s));
    tau_merge_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, s>>>(slist, nq, pl.sParts * SAMPLE_WG, pl.sM, tau);
    NWK_LAUNCH_CHECK();
  } else {
    fill_f32_kernel<<<64, 256, 0, s>>>(tau, pl.nq_pad, HUGE_VALF);
    NWK_LAUNCH_CHECK();
  }
  FinParams f{};
  f.base = base; f.query = query; f.nb = nb; f.nq = nq; f.d = d; f.k = k; f.metric = metric;
  f.id_offset = id_offset; f.tau = tau; f.cnt = cnt; f.log = log; f.sc = sc; f.f16 = f16 ? 1 : 0;
  f.rr = rr; f.rr2 = rr2; f.epsq = epsq; f.qn2 = qn2; f.direct = direct ? 1 : 0; f.alpha = alpha;
  f.bflag = bound ? bflag : nullptr;
  // Relative error of acc against sum |q_i b_i| <= |q||b|:
  //   fp16 (round-to-nearest, 11-bit significand):  (1 + 2^-11)^2 - 1 per product
  //   tf32 (the tensor core drops the low 13 bits): (1 + 2^-10)^2 - 1 per product
  // plus the fp32 accumulation inside the tensor core, budgeted at 4 ulp per MMA step.
  const double steps = (double)d / (f16 ? 16.0 : 8.0) + 2.0;
  f.c_rel = (float)((f16 ? 1.0 / 1024.0 : 1.0 / 512.0) * 1.001 + steps / 4194304.0);
  f.c_acc = (float)((steps + 1.0) / 4194304.0);
  f.out_dist = out_dist; f.out_idx = out_idx; f.fail_list = fail_list;
  if (bound) {
    // key(D) + eps needs max |b|^2 over the whole slab
    if (prep_aside) NWK_CUDA(cudaStreamWaitEvent(s, ds->ev_chunk[nparts - 1], 0));
    tau_bound_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, s>>>(f, bound, bound_stride, tau, bflag);
    NWK_LAUNCH_CHECK();
  }
  if (profile) cudaEventRecord(ev[2], s);

  // ---- filter launches: one per part, alternating between the caller's stream and a side stream.  The
  // launches are independent of each other (log slots are reserved with atomics), so the CTAs of part
  // p+1 move onto the SMs as the CTAs of part p leave them: no tail between the parts.
  cudaStream_t sf = nparts > 1 ? ds->side[1] : s;
  if (nparts > 1) {
    NWK_CUDA(cudaEventRecord(ds->ev_fork, s));          // thresholds are ready
    NWK_CUDA(cudaStreamWaitEvent(sf, ds->ev_fork, 0));
    if (gate) {
      NWK_CUDA(cudaStreamWaitEvent(sp, ds->ev_fork, 0));
      for (int pi = 1; pi < nparts; ++pi) NWK_TRY(launch_prep(pi));
    }
  }
  for (int pi = 0; pi < nparts; ++pi) {
    cudaStream_t sl = (pi & 1) ? sf : s;
    if (prep_aside) NWK_CUDA(cudaStreamWaitEvent(sl, ds->ev_chunk[pi], 0));
    TcParams pp = p;
    pp.t0 = part_t0[pi];
    pp.nT = part_t0[pi + 1] - part_t0[pi];
    if (pp.nT <= 0) continue;
    pp.nQT = nQTf;
    pp.G = G;
    pp.nChunks = (pp.nT + G - 1) / G;
    pp.nUnits = pp.nChunks * pp.nQT;
    if (pair) {
      const int clusters = workers;
      CUtensorMap tm_bh;
      NWK_TRY(make_tmap(&tm_bh, bh, nb, pl.kcols, true, TB / 2, pl.pitch));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2u * (unsigned)(pp.nUnits < clusters ? pp.nUnits : clusters));
      cfg.blockDim = dim3(TC_THREADS);
      cfg.dynamicSmemBytes = TC_SMEM_BYTES;
      cfg.stream = sl;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (direct) NWK_CUDA(cudaLaunchKernelEx(&cfg, tc_kernel<MODE_FILTER, true, true, true>, tm_q, tm_bh, pp));
      else NWK_CUDA(cudaLaunchKernelEx(&cfg, tc_kernel<MODE_FILTER, true, true>, tm_q, tm_bh, pp));
      count_launch();
    } else if (direct || ares) {
      int grid = pp.nUnits < sms ? pp.nUnits : sms;
      if (direct && ares) tc_kernel<MODE_FILTER, true, false, true, true><<<grid, TC_THREADS_WIDE, TC_SMEM_BYTES, sl>>>(tm_q, tm_b, pp);
      else if (direct) tc_kernel<MODE_FILTER, true, false, true><<<grid, TC_THREADS_WIDE, TC_SMEM_BYTES, sl>>>(tm_q, tm_b, pp);
      else tc_kernel<MODE_FILTER, true, false, false, true><<<grid, TC_THREADS, TC_SMEM_BYTES, sl>>>(tm_q, tm_b, pp);
      NWK_LAUNCH_CHECK();
    } else {
      int grid = pp.nUnits < sms ? pp.nUnits : sms;
      NWK_TRY(f16 ? launch_tc<true>(MODE_FILTER, grid, tm_q, tm_b, pp, sl) : launch_tc<false>(MODE_FILTER, grid, tm_q, tm_b, pp, sl));
    }
  }
  if (nparts > 1) {
    // join: the finish stage needs every part's log entries
    NWK_CUDA(cudaEventRecord(ds->ev_join[1], sf));
    NWK_CUDA(cudaStreamWaitEvent(s, ds->ev_join[1], 0));
  }
  if (profile) cudaEventRecord(ev[3], s);
  if (d_prof) {
    std::vector<unsigned long long> hp((size_t)sms * 4);
    NWK_CUDA(cudaMemcpyAsync(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost, s));
    NWK_CUDA(cudaStreamSynchronize(s));
    double a = 0, o = 0, t = 0, n = 0;
    double is = 0;
    for (int i = 0; i < sms; ++i) { a += hp[i * 4]; o += hp[i * 4 + 1] & 0xffffffffull; is += hp[i * 4 + 1] >> 32; t += hp[i * 4 + 2]; n += hp[i * 4 + 3]; }
    if (n > 0)
      fprintf(stderr, "[nwk prof] MMA issuer per tile: %.0f clk total, %.0f waiting for the accumulator stage, %.0f waiting for operands, %.0f issuing MMAs + commit (%.0f tiles)\n",
              t / n, a / n, o / n, is / n, n);
    cudaFree(d_prof);
  }

  {
    select_kernel<<<(unsigned)((nq + SEL_THREADS / 32 - 1) / (SEL_THREADS / 32)), SEL_THREADS, 0, s>>>(f);
    NWK_LAUNCH_CHECK();
    const int rr_smem = (d + 3) / 4 * 16;
    DevState* ds = dev_state();
    if (rr_smem > ds->rr_optin) {
      NWK_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rr_smem));
      ds->rr_optin = rr_smem;
    }
    // a query's batches are dealt over grid.y CTAs; every warp prefetches one batch ahead
    // (measured best split: 1 at d = 128, 4 at d = 768, 8 at d = 1536 -- longer rows want more
    // CTAs per query in flight)
    long long ysplit = std::max<long long>(((long long)sms * 8 + nq - 1) / nq, (d + 96) / 192);
    dim3 rgrid((unsigned)nq, (unsigned)(ysplit < 1 ? 1 : (ysplit > 8 ? 8 : ysplit)));
    rerank_kernel<<<rgrid, RR_THREADS, rr_smem, s>>>(f);
    NWK_LAUNCH_CHECK();
    output_kernel<<<(unsigned)nq, OUT_THREADS, 0, s>>>(f);
    NWK_LAUNCH_CHECK();
  }
  if (profile) cudaEventRecord(ev[4], s);

  Scalars h{};
  NWK_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
  NWK_CUDA(cudaStreamSynchronize(s));
  if (h.err != 0) return fail(NWK_ERR_INTERNAL, "tcgen05 pipeline watchdog expired (role code %d)", h.err);
  if (h.overflow) {
    // direct squared-L2 keys: some |b|^2 does not fit the scale taken from the queries (base rows
    // more than ~11 x longer than the longest query): repeat the call with alpha / beta keys
    if (profile) for (int i = 0; i < 6; ++i) cudaEventDestroy(ev[i]);
    const long long first_pass = g_launches.load() - launches0;
    const int rc = tc_knn(base, nb, query, nq, d, k, metric, id_offset, out_dist, out_idx, ws, ws_bytes,
                          flags | NWK_FLAG_ALPHA_BETA, stats, s, bound, bound_stride);
    if (rc == NWK_OK && stats) stats->gpu_launches += (int)first_pass;
    return rc;
  }
  float err_ratio;
  memcpy(&err_ratio, &h.err_ratio, 4);
  if (err_ratio > 1.0f)
    return fail(NWK_ERR_INTERNAL, "candidate keys exceeded the proven error bound (ratio %.3f)", err_ratio);

  int n_retried = 0, n_scanned = h.fail_cnt;
  nwk_knn_stats st2{};
  if (h.fail_cnt > 0 && !(flags & NWK_FLAG_NO_FALLBACK) && p.dbg == 0) {
    // Uncertified queries.  Where the first pass used a low threshold rank they get a second chance first:
    // the same pipeline for just these queries (one or a few query tiles: a pass over the fp16 base at HBM
    // speed instead of an fp64 scan of the fp32 base per 8 queries), with the threshold taken at rank SAMPLE_M.
    long long scan_from = 0;
    if (!retry && pl.retry_ok && pl.sample && !getenv("NWK_NO_RETRY")) {
      n_retried = (int)std::min<long long>(h.fail_cnt, RETRY_MAX);
      scan_from = n_retried;
    }
    // fp64 scan, in rounds that fit the provisioned scratch (first: the second chance below reuses this
    // call's query-dependent buffers, the fail list among them)
    long long done = scan_from;
    while (done < h.fail_cnt) {
      long long n = h.fail_cnt - done;
      while (n > 8 && exact_scan_workspace_bytes(nb, n, k) > scan_ws_bytes) n = (n + 1) / 2;
      if (exact_scan_workspace_bytes(nb, n, k) > scan_ws_bytes)
        return fail(NWK_ERR_WORKSPACE, "workspace too small for the exact fallback");
      NWK_TRY(exact_scan(base, nb, query, d, k, metric, id_offset, fail_list + done, n, out_dist, out_idx,
                         scan_ws, scan_ws_bytes, s));
      done += n;
    }
    n_scanned = (int)(h.fail_cnt - scan_from);
    if (n_retried > 0) {
      retry_gather_kernel<<<(n_retried + 7) / 8, 256, 0, s>>>(query, fail_list, n_retried, d, bound, bound_stride, wa.rq,
                                                             wa.rbound, wa.rlist);
      NWK_LAUNCH_CHECK();
      NWK_TRY(tc_knn(base, nb, wa.rq, n_retried, d, k, metric, id_offset, wa.rout_d, wa.rout_i, ws, ws_bytes,
                     (flags | FLAG_INTERNAL_RETRY) & ~NWK_FLAG_PROFILE, &st2, s, bound ? wa.rbound : nullptr, 1));
      retry_scatter_kernel<<<std::min((n_retried * k + 255) / 256, sms * 8), 256, 0, s>>>(wa.rout_d, wa.rout_i, wa.rlist,
                                                                                       n_retried, k, out_dist, out_idx);
      NWK_LAUNCH_CHECK();
      n_scanned += st2.n_fallback_queries;
      if (st2.max_err_ratio > err_ratio) err_ratio = st2.max_err_ratio;
    }
  }
  if (profile) cudaEventRecord(ev[5], s);
  NWK_CUDA(cudaStreamSynchronize(s));
  if (stats) {
    stats->path = f16 ? 1 : 2;
    stats->n_fallback_queries = n_scanned;      // queries that took the fp64 scan
    stats->n_retry_queries = n_retried;         // queries that got a second tensor-core pass (some of them scanned after it)
    stats->max_candidates = (int)h.stat_max[0];
    stats->max_rerank = (int)h.stat_max[1];
    stats->total_candidates = (long long)h.stat_sum[0];
    stats->total_rerank = (long long)h.stat_sum[1];
    stats->gpu_launches = (int)(g_launches.load() - launches0);
    stats->max_err_ratio = err_ratio;
    if (profile) {
      cudaEventElapsedTime(&stats->ms_prep, ev[0], ev[1]);
      cudaEventElapsedTime(&stats->ms_sample, ev[1], ev[2]);
      cudaEventElapsedTime(&stats->ms_filter, ev[2], ev[3]);
      cudaEventElapsedTime(&stats->ms_finish, ev[3], ev[4]);
      cudaEventElapsedTime(&stats->ms_fallback, ev[4], ev[5]);
      cudaEventElapsedTime(&stats->ms_total, ev[0], ev[5]);
    }
  }
  if (profile) for (int i = 0; i < 6; ++i) cudaEventDestroy(ev[i]);
  return NWK_OK;
}

// MODE_DUMP writes raw accumulators; with fp16 operands acc = 2^(sq+sb_j) * dot, and the DOT
// metric's beta_j is exactly -2^-(sq+sb_j), so out * (-beta_j) is the plain dot product.
__global__ void dump_unscale_kernel(float* out, const float* beta, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < TQ * TB) out[i] = out[i] * (-beta[(long long)t * TB + (i % TB)]);
}

int tc_debug_tile(const float* base, long long nb, const float* query, long long nq, int d, long long b0,
                  long long q0, float* out, int use_f16, cudaStream_t s) {
  if (d % 4 != 0 || b0 % TB != 0 || q0 % TQ != 0) return fail(NWK_ERR_ARG, "debug tile: bad alignment");
  NWK_TRY(tc_optin());
  const int sms = sm_count();
  const bool f16 = use_f16 != 0;
  TcPlan pl = make_plan(nb, nq, d, 1, sms, f16);
  float *alpha = nullptr, *beta = nullptr;
  __half *bh = nullptr, *qh = nullptr;
  Scalars* sc = nullptr;
  NWK_CUDA(cudaMalloc(&sc, sizeof(Scalars)));
  NWK_CUDA(cudaMemsetAsync(sc, 0, sizeof(Scalars), s));
  NWK_CUDA(cudaMalloc(&alpha, pl.nb_pad * 4));
  NWK_CUDA(cudaMalloc(&beta, pl.nb_pad * 4));
  CUtensorMap tm_q, tm_b;
  int rc = NWK_OK;
  SampleOut no_sample{};
  no_sample.stride_log2 = -1;
  if (f16) {
    NWK_CUDA(cudaMalloc(&bh, (size_t)nb * pl.dh * 2));
    NWK_CUDA(cudaMalloc(&qh, (size_t)nq * pl.dh * 2));
    qmax_kernel<<<sms, 256, 0, s>>>(query, nq * d, &sc->qabsmax);
    NWK_LAUNCH_CHECK();
    qprep_kernel<<<sms * 2, 256, 0, s>>>(query, nq, d, pl.dh, pl.dh, NWK_METRIC_DOT, 0, sc, qh);
    NWK_LAUNCH_CHECK();
    baseprep_kernel<true><<<sms * 4, 256, 0, s>>>(base, nb, 0, pl.nb_pad, d, pl.dh, NWK_METRIC_DOT, alpha, beta, sc, bh, no_sample, 0, pl.dh);
    NWK_LAUNCH_CHECK();
    rc = make_tmap(&tm_q, qh, nq, pl.dh, true, TQ);
    if (rc == NWK_OK) rc = make_tmap(&tm_b, bh, nb, pl.dh, true, TB);
  } else {
    baseprep_kernel<false><<<sms * 4, 256, 0, s>>>(base, nb, 0, pl.nb_pad, d, pl.dh, NWK_METRIC_DOT, alpha, beta, sc, bh, no_sample, 0, pl.dh);
    NWK_LAUNCH_CHECK();
    rc = make_tmap(&tm_q, query, nq, d, false, TQ);
    if (rc == NWK_OK) rc = make_tmap(&tm_b, base, nb, d, false, TB);
  }
  int h = 0;
  if (rc == NWK_OK) {
    TcParams p{};
    p.nKB = pl.nKB; p.kLast = pl.kLast; p.nq = nq; p.nb = nb; p.nQT = pl.nQT; p.nT = pl.nT;
    p.dump = out; p.dump_qt = (int)(q0 / TQ); p.dump_t = (int)(b0 / TB); p.err = &sc->err;
    rc = f16 ? launch_tc<true>(MODE_DUMP, 1, tm_q, tm_b, p, s) : launch_tc<false>(MODE_DUMP, 1, tm_q, tm_b, p, s);
    if (rc == NWK_OK) {
      dump_unscale_kernel<<<(TQ * TB + 255) / 256, 256, 0, s>>>(out, beta, p.dump_t);
      count_launch();
      Scalars hs{};
      cudaMemcpyAsync(&hs, sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s);
      cudaError_t e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) rc = fail(NWK_ERR_CUDA, "debug tile failed: %s", cudaGetErrorString(e));
      h = hs.err;
    }
  }
  cudaFree(sc); cudaFree(alpha); cudaFree(beta); cudaFree(bh); cudaFree(qh);
  if (rc != NWK_OK) return rc;
  if (h != 0) return fail(NWK_ERR_INTERNAL, "tcgen05 pipeline watchdog expired (role code %d)", h);
  return NWK_OK;
}

#include "nwk_maxsim.cuh"

}  // namespace nwk
