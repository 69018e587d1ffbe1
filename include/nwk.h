/* nwk.h -- C ABI of the B200-native exact brute-force kNN engine (libnwk.so).
 *
 * This is the drop-in boundary for neighborhoodwatch's ground-truth hot path.  The
 * reference has no FFI layer of its own (it is pure Python calling RAPIDS / torch), so
 * every entry point below replaces a library call site or a Python function of the
 * reference; the file:line each one stands in for is cited beside it (paths relative to
 * /root/reference/).  INTEGRATION.md shows the ctypes stubs a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - C linkage, plain pointers and sizes only; no C++ or torch types cross the boundary.
 *   - Every function returns 0 on success and a non-zero nwk_status otherwise;
 *     nwk_last_error() returns a thread-local human-readable message for the last failure.
 *   - All matrix arguments are dense row-major (C order) fp32, the layout the reference
 *     produces with cupy ``.copy(order='C')`` (neighborhoodwatch/cu_knn.py:211,235).
 *   - Unless a parameter is documented as HOST memory, pointers are DEVICE addresses valid
 *     on the current CUDA device.  The caller owns every buffer; the library allocates only
 *     inside opaque handles (nwk_topk_state) that it also frees.
 *   - ``stream`` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *     Calls enqueue work on that stream; nwk_knn_f32 additionally synchronises the stream
 *     before returning (the reference is synchronous too: every knn() is followed by
 *     ``.get()``, cu_knn.py:271) because it must read back the per-query certification
 *     counters to decide whether the exact fallback pass is needed.
 *   - There is no CPU fallback: on a machine without an sm_100 device every compute entry
 *     point fails with NWK_ERR_DEVICE.
 */
#ifndef NWK_H_
#define NWK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NWK_ABI_VERSION 2

typedef enum nwk_status {
  NWK_OK = 0,
  NWK_ERR_ARG = 1,        /* bad argument (null pointer, k > rows, d == 0, ...)       */
  NWK_ERR_CUDA = 2,       /* a CUDA runtime / driver call failed                      */
  NWK_ERR_WORKSPACE = 3,  /* workspace smaller than nwk_knn_workspace_bytes()         */
  NWK_ERR_DEVICE = 4,     /* no sm_100 device, or the kernel image cannot run here    */
  NWK_ERR_INTERNAL = 5    /* an internal invariant failed (kernel watchdog, overflow) */
} nwk_status;

/* Distance definitions (selected in the reference by the ``engine`` argument of
 * process_batches, cu_knn.py:242-262):
 *   NWK_METRIC_SQEUCLIDEAN  sum_i (q_i - b_i)^2          engine 'raft'  cu_knn.py:243-245
 *   NWK_METRIC_COSINE       1 - q.b / (|q| |b|)          engine 'cuvs'  cu_knn.py:247-248
 *   NWK_METRIC_DOT          1 - q.b                      engine 'torch' cu_knn.py:256     */
typedef enum nwk_metric {
  NWK_METRIC_SQEUCLIDEAN = 0,
  NWK_METRIC_COSINE = 1,
  NWK_METRIC_DOT = 2
} nwk_metric;

/* flags for nwk_knn_f32 / nwk_knn_host_f32 */
#define NWK_FLAG_FORCE_EXACT_SCAN 0x1u /* skip the tensor-core candidate stage; fp64 scan of every pair */
#define NWK_FLAG_NO_FALLBACK      0x2u /* do not run the fp64 scan for uncertified queries (diagnostics) */
#define NWK_FLAG_PROFILE          0x4u /* bracket each stage with CUDA events on `stream`; fills stats->ms_*  */
#define NWK_FLAG_TF32             0x8u /* candidate stage on fp32 operands as TF32 (no fp16 staging copy)      */
#define NWK_FLAG_ALPHA_BETA      0x10u /* candidate keys as alpha_j + beta_j * acc for every metric (no direct keys) */

/* Per-call diagnostics (all counters refer to the last nwk_knn_f32 call they were passed to). */
typedef struct nwk_knn_stats {
  int32_t path;               /* 0 = fp64 exact scan; tcgen05 candidates + fp64 rerank: 1 = fp16, 2 = TF32; 3 = nwk_maxsim_f32 */
  int32_t n_fallback_queries; /* queries whose candidate set could not be certified (after the second chance) -> fp64 exact scan */
  int32_t max_candidates;     /* largest per-query candidate-log population seen                         */
  int32_t max_rerank;         /* largest per-query number of rows re-ranked in fp64                      */
  int64_t total_candidates;   /* sum over queries of candidate-log population                            */
  int64_t total_rerank;       /* sum over queries of rows re-ranked in fp64                              */
  int32_t gpu_launches;       /* kernels launched by this call                                           */
  float   max_err_ratio;      /* max |candidate key - exact key| / proven bound over re-ranked rows (<= 1) */
  /* device time per stage, CUDA events on the caller's stream; 0 unless NWK_FLAG_PROFILE */
  float   ms_prep;            /* operand preparation: norms, alpha/beta, fp16 staging copy (HBM-bound)   */
  float   ms_sample;          /* tcgen05 sample pass -> per-query thresholds                             */
  float   ms_filter;          /* tcgen05 filter pass over the whole base: the dominant kernel            */
  float   ms_finish;          /* select + fp64 rerank                                                    */
  float   ms_fallback;        /* fp64 scan of uncertified queries                                        */
  float   ms_total;           /* whole call                                                              */
  int32_t n_retry_queries;    /* uncertified queries that got a second tensor-core pass at a higher threshold */
  int32_t reserved_;
} nwk_knn_stats;

const char* nwk_last_error(void);
int nwk_abi_version(void);

/* Number of kernels this library has launched in the calling process (bench.py's
 * ``gpu_launches`` is the difference of two readings). */
int64_t nwk_launch_count(void);

/* Exact k nearest neighbours of every query row among the base rows.
 *
 * Replaces the array-level operator of the reference:
 *   pylibraft.neighbors.brute_force.knn(dataset, queries, k)   cu_knn.py:243-245, cu_knn_ds.py:213
 *   cuvs.neighbors.brute_force.build(...)/search(...)          cu_knn.py:247-248
 *   1 - torch.matmul(q, b.T); torch.topk(.., largest=False)    cu_knn.py:256-257
 * and folds in the ``indices + batch_offset`` step of cu_knn.py:273 through ``id_offset``.
 *
 *   base      [nb, d] fp32 row-major, device
 *   query     [nq, d] fp32 row-major, device
 *   out_dist  [nq, k] fp32, ascending per row
 *   out_idx   [nq, k] int64, row index into base + id_offset
 *   workspace device scratch of at least nwk_knn_workspace_bytes(...) bytes, 256-byte aligned
 *   stats     optional (may be NULL), HOST memory
 *
 * Results are exact: candidates are produced on the tensor cores (fp16 operands scaled by
 * exact powers of two -- the significand width of TF32 -- or TF32 itself), every candidate
 * within a proven error bound of the k-th is re-ranked in fp64, and queries for which the
 * bound cannot be established are recomputed by an fp64 scan.  Ordering is by
 * (fp64 distance, index).  base / query pointers that are not 16-byte aligned are served by
 * the fp64 scan (correct, slow).  Fails with NWK_ERR_ARG when k > nb (the reference asserts
 * ``k <= len(dataset)``, cu_knn.py:237) or k > NWK_MAX_K. */
#define NWK_MAX_K 1024
size_t nwk_knn_workspace_bytes(int64_t nb, int64_t nq, int32_t d, int32_t k, int32_t metric,
                               uint32_t flags);
int nwk_knn_f32(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d,
                int32_t k, int32_t metric, int64_t id_offset, float* out_dist, int64_t* out_idx,
                void* workspace, size_t ws_bytes, uint32_t flags, nwk_knn_stats* stats,
                void* stream);

/* Merge P sorted partial results into the k_out best per query, ordered by
 * (distance, index).  Replaces merge_indices_and_distances' argsort fold, merge.py:73-104,
 * and is the cross-GPU / cross-slab reducer.
 *
 *   dist  [P, nq, k_in] fp32 device (partial p, query q, slot j at ((p*nq)+q)*k_in + j)
 *   idx   [P, nq, k_in] int64 device, global row ids
 *   out_dist [nq, k_out] fp32, out_idx [nq, k_out] int64; requires k_out <= P*k_in
 *   workspace: nwk_merge_workspace_bytes(P, nq, k_in, k_out) bytes of device scratch */
#define NWK_MERGE_MAX_K 2048
size_t nwk_merge_workspace_bytes(int32_t P, int64_t nq, int32_t k_in, int32_t k_out);
int nwk_merge_topk(int32_t P, const float* dist, const int64_t* idx, int64_t nq, int32_t k_in,
                   int32_t k_out, float* out_dist, int64_t* out_idx, void* workspace,
                   size_t ws_bytes, void* stream);
/* The same merge when the P lists are not one contiguous [P, nq, k_in] array.
 *   _strided: list p is dist + p * dist_stride / idx + p * idx_stride (strides in ELEMENTS) -- the
 *             chunks of a packed all-gather buffer [P][nq*k_in f32 | nq*k_in i64], so one
 *             collective moves distances and ids together;
 *   _ptrs:    list p is dist_ptrs[p] / idx_ptrs[p] (HOST arrays of P <= 16 device pointers): lists
 *             that live in different allocations or in the memory of peer GPUs (the kernel then
 *             reads them over NVLink: gather and merge are one kernel).
 * Padding entries are those with a negative id (they come out as +inf / -1 again); entries with a
 * real id keep it even when their distance is +inf or nan (ordered last).  k_out <= NWK_MERGE_MAX_K. */
int nwk_merge_topk_strided(int32_t P, const float* dist, int64_t dist_stride, const int64_t* idx,
                           int64_t idx_stride, int64_t nq, int32_t k_in, int32_t k_out, float* out_dist,
                           int64_t* out_idx, void* workspace, size_t ws_bytes, void* stream);
int nwk_merge_topk_ptrs(int32_t P, const float* const* dist_ptrs, const int64_t* const* idx_ptrs, int64_t nq,
                        int32_t k_in, int32_t k_out, float* out_dist, int64_t* out_idx, void* workspace,
                        size_t ws_bytes, void* stream);

/* Running per-query top-k across base slabs: the intended semantics of the streaming
 * driver, cu_knn_ds.py:181-239 (one global result over all base batches), and the
 * reducer compute_knn uses instead of writing P partial files when asked to.
 * update() merges one [nq, k] partial (device pointers) into the state; finalize() copies
 * the current result out.  The handle owns its device memory. */
typedef struct nwk_topk_state nwk_topk_state;
int nwk_topk_state_create(int64_t nq, int32_t k, nwk_topk_state** out);
int nwk_topk_state_update(nwk_topk_state* st, const float* dist, const int64_t* idx, void* stream);
/* search() = "knn of one more slab of the base, folded into the state" in one call: what the
 * loop body of cu_knn_ds.py:201-226 does with knn() + concat, made incremental.  The state's
 * current k-th distances bound the search (a slab row can only enter the result if it beats the
 * query's k-th best so far), so slabs after the first log and re-rank a fraction of the
 * candidates; results are identical to update(knn(slab)).  base / query / workspace as for
 * nwk_knn_f32 with nq and k taken from the state (nb >= k, workspace of
 * nwk_knn_workspace_bytes(nb, nq, d, k, ...)); synchronises the stream like nwk_knn_f32. */
int nwk_topk_state_search(nwk_topk_state* st, const float* base, int64_t nb, const float* query, int32_t d,
                          int32_t metric, int64_t id_offset, void* workspace, size_t ws_bytes, uint32_t flags,
                          nwk_knn_stats* stats, void* stream);
/* finalize() copies the current result to out_dist / out_idx (device or host memory, enqueued on
 * `stream`); reset() empties the state on `stream` so the handle can serve the next run. */
int nwk_topk_state_finalize(nwk_topk_state* st, float* out_dist, int64_t* out_idx, void* stream);
int nwk_topk_state_reset(nwk_topk_state* st, void* stream);
int nwk_topk_state_destroy(nwk_topk_state* st);

/* Host-buffer entry point (what `e2e` in bench.py times): base and query are HOST
 * pointers (pinned memory gives full PCIe rate); the base is cut into slabs of
 * ``slab_rows`` rows that are copied on a side stream into a two-deep device ring while
 * the previous slab is searched; results are written to HOST out_dist / out_idx.
 * Replaces the load + loop of compute_knn_ds / process_dataset_batches
 * (cu_knn_ds.py:136-239) and of compute_knn / process_batches (cu_knn.py:151-287) for
 * callers that hold the vectors in host memory.  slab_rows == 0 picks a default. */
int nwk_knn_host_f32(const float* base_host, int64_t nb, const float* query_host, int64_t nq,
                     int32_t d, int32_t k, int32_t metric, int64_t id_offset, int64_t slab_rows,
                     float* out_dist_host, int64_t* out_idx_host, uint32_t flags,
                     nwk_knn_stats* stats);
/* The slab partition nwk_knn_host_f32 uses (pure arithmetic, no device needed): slabs 0 .. nslab-2
 * hold *slab_rows rows, the last one *last_rows; every slab holds at least k rows (the reference
 * asserts the same of its last batch, cu_knn.py:177 / cu_knn_ds.py:168 -- here a tail shorter than k
 * is folded into the slab before it).  slab_rows_in == 0 picks the default. */
int nwk_host_slab_plan(int64_t nb, int64_t nq, int32_t d, int32_t k, int64_t slab_rows_in,
                       int64_t* slab_rows, int64_t* nslab, int64_t* last_rows);
/* Pinned (page-locked) host memory for the arrays nwk_knn_host_f32 streams from; write_combined != 0 maps
 * it write-combined (fast for DMA, slow for CPU reads).  Free with nwk_host_free. */
int nwk_host_alloc(size_t bytes, int32_t write_combined, void** out);
int nwk_host_free(void* p);
/* The host path keeps its device staging ring, workspace and running state per device between
 * calls (no allocation once they have grown to the shapes in use); this frees them. */
int nwk_host_release(void);

/* Cross-GPU reduction over peer memory: gather + k-way merge as ONE kernel (csrc/nwk_xchg.cu).
 * One process per GPU on one NVLink / NVSwitch box.  Every rank creates a handle for the same
 * (world, nq, k); the library allocates the rank's exported result buffer inside it.  The ranks
 * swap the opaque handles of nwk_xchg_local_handle (nwk_xchg_handle_bytes() each, any transport:
 * the Python driver uses one torch.distributed all-gather at start-up) and map them with
 * nwk_xchg_open_peers (CUDA IPC).  Per step: write the local top-k (nwk_knn_f32 with out_dist /
 * out_idx = the pointers of nwk_xchg_next_slot, ids global), then nwk_xchg_merge: it publishes the
 * rank's step counter; the rank's merge kernel waits -- on the device, bounded -- until every peer
 * has published the same step, merges ITS SLICE of the queries (nq / world of them) reading the
 * world lists of each query straight out of the owners' memory over NVLink, and stores the merged
 * rows into every rank's result buffer; once all ranks have done so the [nq, k_out] result
 * (ordered by (distance, id), padding = negative ids, as nwk_merge_topk; k_out <= k) is copied to
 * out_dist / out_idx.  Every record crosses the fabric once (an all-gather moves it world - 1
 * times); no collective library call, no staging copy and no host round trip on the step path.
 * All ranks must call nwk_xchg_merge the same number of times.  Replaces, across GPUs, what
 * merge_indices_and_distances (merge.py:56-101) does across the reference's base batches.
 * nwk_xchg_check reports a timed-out wait (a rank that never published). */
typedef struct nwk_xchg nwk_xchg;
int nwk_xchg_create(int32_t rank, int32_t world, int64_t nq, int32_t k, nwk_xchg** out);
size_t nwk_xchg_handle_bytes(void);
int nwk_xchg_local_handle(nwk_xchg* x, void* handle_out);
int nwk_xchg_open_peers(nwk_xchg* x, const void* handles);
int nwk_xchg_next_slot(nwk_xchg* x, float** dist, int64_t** idx);
int nwk_xchg_merge(nwk_xchg* x, int32_t k_out, float* out_dist, int64_t* out_idx, void* stream);
int nwk_xchg_check(nwk_xchg* x);
int nwk_xchg_destroy(nwk_xchg* x);

/* Single-process multi-GPU (SURVEY.md section 8b's nwk_mg_*): one host process, n devices of one box.
 * Shard i of the base (rows[i] >= k rows, row-major fp32) lives on devices[i]; the queries are HOST memory
 * and are replicated by the call; ids are global (id_offset + the rows of the shards before).  Each device
 * runs nwk_knn_f32 on a host thread of its own; device 0 then merges the n lists reading the peers' lists
 * in place over NVLink (peer access: gather + merge are one kernel) and the [nq, k] result lands in HOST
 * out_dist / out_idx.  The handle owns its per-device buffers; n * k <= 2048.  Replaces the reference's
 * temporal batch slicing + CPU merge (cu_knn.py:201-205, merge.py:56-101) for callers without a launcher;
 * with one process per GPU use nwk_xchg_*. */
typedef struct nwk_mg nwk_mg;
int nwk_mg_create(const int32_t* devices, int32_t n, nwk_mg** out);
int nwk_mg_knn_f32(nwk_mg* mg, const float* const* base_shards, const int64_t* rows, const float* query_host,
                   int64_t nq, int32_t d, int32_t k, int32_t metric, int64_t id_offset, float* out_dist_host,
                   int64_t* out_idx_host, uint32_t flags, nwk_knn_stats* stats);
int nwk_mg_destroy(nwk_mg* mg);

/* Document-level ColBERT MaxSim (the reference's colbert_knn.py:62-126 stops at a
 * token-level kNN over the flattened token embeddings, which nwk_knn_f32 with NWK_METRIC_DOT covers):
 *
 *   score(Q, D) = sum over Q's tokens t of  max over D's tokens j of  q_t . d_j,   top-k documents per query.
 *
 *   query_tokens [nq, tq, d] fp32 device (tq a power of two <= 128; pad shorter queries with zero tokens,
 *                 which add 0 to every score), doc_tokens [nd, td, d] fp32 device (td <= 256),
 *   doc_len      [nd] int32 device or NULL: valid tokens of each document (the rest is padding and ignored),
 *   out_score    [nq, k] fp32 descending, out_idx [nq, k] int64 = document number + id_offset
 *                (ties by document number); k <= min(256, nd); d % 4 == 0.
 * Approximate scores come from a tcgen05 GEMM with the max / sum fused into its TMEM epilogue; the
 * 4k best documents per query are re-scored exactly in fp64 and the result is certified against a
 * proven error bound (uncertified queries are re-scored against all documents), like nwk_knn_f32.
 * stats->path = 3.  Synchronises the stream. */
size_t nwk_maxsim_workspace_bytes(int64_t nq, int32_t tq, int64_t nd, int32_t td, int32_t d, int32_t k);
int nwk_maxsim_f32(const float* query_tokens, int64_t nq, int32_t tq, const float* doc_tokens, int64_t nd, int32_t td,
                   const int32_t* doc_len, int32_t d, int32_t k, int64_t id_offset, float* out_score, int64_t* out_idx,
                   void* workspace, size_t ws_bytes, nwk_knn_stats* stats, void* stream);

/* Deterministic synthetic vectors (the fixture style of tests/test_02_cu_knn.py:21-28:
 * uniform[0,1) components, rows L2-normalised), counter-based so any slab of a large base
 * can be regenerated: element (row, col) depends only on (seed, row, col, d).  The numpy
 * twin is neighborhoodwatch_b200/synth.py.  out is [rows, d] fp32 device memory holding
 * rows row0 .. row0+rows-1. */
int nwk_synth_fill(float* out, int64_t rows, int32_t d, uint64_t seed, int64_t row0,
                   int32_t normalize, void* stream);

/* Column-per-dimension staging -> the row-major array knn() takes: src is [d, ld] fp32 (element
 * (row r, dimension c) at src[c * ld + r], the layout of the reference's parquet files,
 * generate_dataset.py:233-235), dst is [rows, d] row-major.  Replaces cupy.copy(order='C') at
 * cu_knn.py:211,235 / cu_knn_ds.py:197,211.  Device pointers, enqueued on `stream`. */
int nwk_rows_from_columns_f32(const float* src, int64_t ld, int32_t d, int64_t rows, float* dst, void* stream);

/* Post-validator with the reference's acceptance rule (parquet_to_format.py:351-383): for every
 * query q and returned neighbour j, numpy.isclose(1 - dot(query_q, base[idx[q,j]]), dist[q,j] / 2)
 * (squared L2 on unit vectors; the reference's defaults are rtol = 1e-5, atol = 1e-8); zero query
 * vectors are skipped as there.  All pointers are device memory except mismatches_host, which
 * receives the number of violating pairs (the call synchronises the stream). */
int nwk_validate_pairs(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d,
                       const int64_t* idx, const float* dist, int32_t k, double rtol, double atol,
                       int64_t* mismatches_host, void* stream);

/* Diagnostics used by the parity tests: raw tensor-core dot products (fp16 or TF32 operands) of one
 * 128 x 256 tile pair (query rows q0.., base rows b0..) written to out[128*256] fp32,
 * so the tcgen05 / TMA descriptors can be validated in isolation. */
int nwk_debug_tc_tile(const float* base, int64_t nb, const float* query, int64_t nq, int32_t d,
                      int64_t b0, int64_t q0, float* out, int32_t use_f16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NWK_H_ */
