// nwk_maxsim.cuh -- document-level ColBERT MaxSim, included at the end of nwk_tc.cu (same namespace,
// same TMA / tcgen05 helpers).
//
//   score(Q, D) = sum over the query's tokens t of  max over the document's tokens j of  q_t . d_j
//
// The reference (colbert_knn.py:62-126) only runs a token-level kNN over the flattened token
// embeddings; SURVEY.md 8(f) lists the document-level score as the next metric.  Pipeline:
//
//   qmax x 2, qprep, ms_docprep   tokens -> fp16 with one power-of-two scale per side (exact scaling); every
//                               document padded to N = ceil(Td / 32) * 32 token rows = one MMA tile
//   maxsim_kernel               tcgen05 GEMM of 128 query-token rows (128 / Tq queries) x one document per tile;
//                               the epilogue takes each row's maximum over the document's valid columns straight
//                               out of TMEM and sums the Tq rows of a query with warp shuffles: one approximate
//                               score per (query, document), the token-level similarity matrix never exists
//   topk_from_keys              the kc best approximate scores per query (segmented bitonic top-k, nwk_scan.cu)
//   ms_exact_kernel             exact fp64 MaxSim of those (query, document) pairs from the fp32 inputs
//   ms_final_kernel             order by (exact score descending, document id), emit k, certify: every document
//                               outside the candidates has an approximate score more than 2 E_q below the k-th
//                               (E_q = proven bound on |approximate - exact|); uncertified queries are recomputed
//                               exactly against all documents.
#pragma once

struct MsParams {
  int nKB, kLast;
  long long nq, nd;
  int Tq, Td, N;              // tokens per query (power of two <= 128), per document (<= 256), tile columns
  const int* doc_len;         // [nd] valid tokens per document, or null (= Td)
  int nQT, G, nUnits;         // query tiles, documents per unit, units = chunks x query tiles
  const Scalars* sc;          // qabsmax / babsmax: the two scales
  double* neg_score;          // [nq][nd]  -(approximate score)
  int* err;
};

constexpr int MS_THREADS = 384;   // 4 control warps + 2 epilogue warpgroups that take the tiles (= accumulator stages) in turn

// documents -> fp16 rows [nd * N][pitch]: token j of document i at row i * N + j, rows j >= Td zero;
// max |d|^2 over all tokens for the error bound.  One warp per output row.
__global__ void __launch_bounds__(256)
ms_docprep_kernel(const float* __restrict__ docs, long long nd, int Td, int N, int d, int pitch, Scalars* __restrict__ sc,
                  __half* __restrict__ dh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = pow2f(f16_scale_exp(__uint_as_float(sc->babsmax)));
  float local_max = 0.f;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < nd * N; r += (long long)gridDim.x * 8) {
    const long long doc = r / N;
    const int j = (int)(r - doc * N);
    __half* out = dh + r * pitch;
    float a = 0.f;
    for (int c = lane * 4; c < pitch; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < Td && c < d) v = __ldg(reinterpret_cast<const float4*>(docs + (doc * Td + j) * d + c));
      a = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, a))));
      __half2 h01 = __floats2half2_rn(v.x * scale, v.y * scale);
      __half2 h23 = __floats2half2_rn(v.z * scale, v.w * scale);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned int*>(&h01);
      pk.y = *reinterpret_cast<unsigned int*>(&h23);
      *reinterpret_cast<uint2*>(out + c) = pk;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (a < __int_as_float(0x7f800000)) local_max = fmaxf(local_max, a * 1.0001f);
  }
  if (lane == 0 && local_max > 0.f) atomicMax(&sc->bn2max, __float_as_uint(local_max));
}

__global__ void __launch_bounds__(MS_THREADS, 1)
maxsim_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d, const MsParams p) {
  constexpr int BK = 64;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [ACC_STAGES]
  uint64_t* tempty_bar = bars + 2 * STAGES + ACC_STAGES;
  uint64_t* afull_bar = bars + 2 * STAGES + 2 * ACC_STAGES;        // [2] resident query tile buffers
  uint64_t* aempty_bar = bars + 2 * STAGES + 2 * ACC_STAGES + 2;   // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The query tile stays resident for a whole unit (G documents), like tc_kernel's ARES variant: the
  // ring then carries document stages only (N x 128 bytes per K-block instead of 16 KB more), which is
  // what keeps a 128 x 192 x 128 tile (768 clk of MMAs) inside the 64 B/clk an SM can take from L2.
  // Rows too long for that (more than 6 K-blocks) stream the query tile with every document.
  const bool ares = p.nKB <= 6;
  const int a_bytes = p.nKB * STAGE_A_BYTES;
  const int nA = ares ? ((2 * a_bytes + 4 * STAGE_B_BYTES <= STAGES * STAGE_BYTES) ? 2 : 1) : 0;
  const int nsb = ares ? min(STAGES, (STAGES * STAGE_BYTES - nA * a_bytes) / STAGE_B_BYTES) : STAGES;
  unsigned char* const b_ring = smem + nA * a_bytes;
  const int stage_bytes = (ares ? 0 : STAGE_A_BYTES) + p.N * 128;

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&tmap_q); ptx::prefetch_tensormap(&tmap_d); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 4);
      ptx::mbar_init(&afull_bar[a], 1); ptx::mbar_init(&aempty_bar[a], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(s_tmem, ACC_STAGES * TB); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // the schedule all roles walk: unit u = (chunk of G documents, query tile), chunk-major, so the CTAs
  // that run together read the same documents (fetched from HBM once, served from L2 to the others)
  auto unit = [&](int u, int& qt, long long& doc0, int& len) {
    const int c = u / p.nQT;
    qt = u - c * p.nQT;
    doc0 = (long long)c * p.G;
    len = (int)min((long long)p.G, p.nd - doc0);
  };

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    int abuf = 0; uint32_t aphase = 0;
    bool ok = true;
    for (int u = blockIdx.x; ok && u < p.nUnits; u += gridDim.x) {
      int qt, len; long long doc0;
      unit(u, qt, doc0, len);
      if (ares) {
        if (!ptx::mbar_wait(&aempty_bar[abuf], aphase ^ 1, p.err, ERR_PRODUCER)) { ok = false; break; }
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&afull_bar[abuf], a_bytes);
          for (int kb = 0; kb < p.nKB; ++kb)
            ptx::tma_load_2d(smem + abuf * a_bytes + kb * STAGE_A_BYTES, &tmap_q, &afull_bar[abuf], kb * BK, qt * TQ);
        }
        __syncwarp();
        if (++abuf == nA) { abuf = 0; aphase ^= 1; }
      }
      for (int i = 0; ok && i < len; ++i)
        for (int kb = 0; kb < p.nKB; ++kb) {
          if (!ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.err, ERR_PRODUCER)) { ok = false; break; }
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            if (ares) {
              ptx::tma_load_2d(b_ring + stage * STAGE_B_BYTES, &tmap_d, &full_bar[stage], kb * BK, (int)((doc0 + i) * p.N));
            } else {
              unsigned char* sa = smem + stage * STAGE_BYTES;
              ptx::tma_load_2d(sa, &tmap_q, &full_bar[stage], kb * BK, qt * TQ);
              ptx::tma_load_2d(sa + STAGE_A_BYTES, &tmap_d, &full_bar[stage], kb * BK, (int)((doc0 + i) * p.N));
            }
          }
          __syncwarp();
          if (++stage == nsb) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t idesc = ptx::umma_idesc_f16(TQ, p.N);
    int abuf = 0; uint32_t aphase = 0;
    bool ok = true;
    for (int u = blockIdx.x; ok && u < p.nUnits; u += gridDim.x) {
      int qt, len; long long doc0;
      unit(u, qt, doc0, len);
      if (ares) {
        if (!ptx::mbar_wait(&afull_bar[abuf], aphase, p.err, ERR_MMA_FULL)) { ok = false; break; }
      }
      for (int i = 0; ok && i < len; ++i) {
        if (!ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err, ERR_MMA_EMPTY)) { ok = false; break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TB;
        for (int kb = 0; kb < p.nKB; ++kb) {
          if (!ptx::mbar_wait(&full_bar[stage], phase, p.err, ERR_MMA_FULL)) { ok = false; break; }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t sa = ares ? ptx::smem_u32(smem + abuf * a_bytes + kb * STAGE_A_BYTES)
                                     : ptx::smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t da = ptx::umma_desc_sw128(sa);
            const uint64_t db = ptx::umma_desc_sw128(ares ? ptx::smem_u32(b_ring + stage * STAGE_B_BYTES) : sa + STAGE_A_BYTES);
            const int nk = (kb == p.nKB - 1) ? p.kLast : BK / 16;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (k >= nk) break;
              ptx::mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::tc_commit(&empty_bar[stage]);
            if (kb == p.nKB - 1) {
              ptx::tc_commit(&tfull_bar[acc]);
              if (ares && i == len - 1) ptx::tc_commit(&aempty_bar[abuf]);   // the unit's MMAs are all issued
            }
          }
          __syncwarp();
          if (++stage == nsb) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      if (ares) { if (++abuf == nA) { abuf = 0; aphase ^= 1; } }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread <-> TMEM lane <-> query-token row of the tile.  Two warpgroups, each
    // bound to one accumulator stage: warpgroup g takes every second tile whole (all N columns), so
    // a row's maximum never has to be combined across warpgroups and each has two MMA tile times
    // for its tile (one warpgroup alone, 6 x (load + 16 dependent FMNMX3), took longer than the
    // 768 clk of a tile's MMAs). =====
    const int quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const float NINF = __int_as_float(0xff800000);
    int ti = 0;                                                     // tiles this CTA has walked
    int sq, sb;
    direct_exps(NWK_METRIC_DOT, p.sc, sq, sb);
    const double unscale = (double)pow2f(-(sq + sb));
    bool ok = true;
    for (int u = blockIdx.x; ok && u < p.nUnits; u += gridDim.x) {
      int qt, len; long long doc0;
      unit(u, qt, doc0, len);
      const long long row = (long long)qt * TQ + row_in_tile;     // query-token row
      const long long q = row / p.Tq;
      for (int i = 0; ok && i < len; ++i) {
        const int acc = ti & 1;
        const uint32_t acc_phase = (ti >> 1) & 1;
        ++ti;
        if (acc != wg) continue;
        const long long doc = doc0 + i;
        if (!ptx::mbar_wait(&tfull_bar[acc], acc_phase, p.err, ERR_EPI)) { ok = false; break; }
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * TB;
        const int valid = p.doc_len ? max(0, min(p.doc_len[doc], p.Td)) : p.Td;    // warp-uniform
        const int nch = (valid + 31) >> 5;
        float m0 = NINF, m1 = NINF;
        uint32_t va[32], vb[32];
        auto fold = [&](const uint32_t (&v)[32], const int left) {
          if (left >= 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < left) m0 = fmaxf(m0, __uint_as_float(v[j]));
          }
        };
        if (nch > 0) ptx::tmem_ld_32x32(taddr, va);
        for (int ch = 0; ch < nch; ch += 2) {
          ptx::tmem_wait_ld();
          if (ch + 1 < nch) ptx::tmem_ld_32x32(taddr + (ch + 1) * 32, vb);
          fold(va, valid - ch * 32);
          if (ch + 1 < nch) {
            ptx::tmem_wait_ld();
            if (ch + 2 < nch) ptx::tmem_ld_32x32(taddr + (ch + 2) * 32, va);
            fold(vb, valid - (ch + 1) * 32);
          }
        }
        float m = fmaxf(m0, m1);
        __syncwarp();
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        // sum over the query's token rows (Tq a power of two): inside the warp, and across warps by atomics
        const int span = min(p.Tq, 32);
        for (int o = span >> 1; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
        if ((lane & (span - 1)) == 0 && q < p.nq) {
          const double val = -(double)m * unscale;
          if (p.Tq <= 32) p.neg_score[q * p.nd + doc] = val;
          else atomicAdd(p.neg_score + q * p.nd + doc, val);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, ACC_STAGES * TB); }
}

// E_q: proven bound on |approximate - exact| score of one query = sum over its tokens of the direct-dot
// bound of key_error_bound() (operand rounding + accumulation, quantisation of tiny elements on both
// sides), plus the fp32 roundings of the epilogue's sum and of the stored key.  sum_norm = sum_t |q_t|.
__device__ __forceinline__ double ms_error_bound(const Scalars* sc, float c_rel, int Tq, int d, double sum_norm) {
  int sq, sb;
  direct_exps(NWK_METRIC_DOT, sc, sq, sb);
  const double dmax = sqrt((double)__uint_as_float(sc->bn2max));
  return ((double)c_rel * dmax + ldexp(1.0, -25 - sb) * sqrt((double)d) + 4e-6 * dmax) * sum_norm +
         Tq * ldexp(1.0, -25 - sq) * sqrt((double)d) * dmax + 1e-30;
}

// exact fp64 MaxSim of (query, document) pairs: grid (pairs per query, queries); lane = query token,
// warp = document tokens j = warp, warp + 4, ... of a tile.
struct MsExact {
  const float* qtok; const float* docs; const int* doc_len;
  long long nq, nd; int Tq, Td, d;
  const int* qsel;            // null: queries 0..; else the selected queries
  const long long* cand;      // [.][nc] document ids per (selected) query, or null = documents 0..nc-1
  long long nc;
  double* out;                // [.][nc]  -(exact score)
  // candidates only: a document whose approximate key lies more than 2 E_q above the k-th best cannot be
  // among the true top k and is not scored (its slot gets +inf)
  const float* cand_key; int k;
  const float* eq;            // [.] E_q of the (selected) queries (ms_eq_kernel)
};

// The document is staged tile by tile (TJ = 4 PW tokens) in shared memory as doubles, converted once for
// all query tokens; lane = query token keeps 8 of its components as doubles in registers per pass and
// one accumulator per document token of its warp (PW of them: tokens warp, warp + 4, ...).  The inner
// loop is straight-line: 4 broadcast LDS.128 + 8 DFMA per document token, no branches (the tokens past
// the end of the document are staged as zeros and left out of the maximum afterwards).
// Round 2 rebalanced it around the fp64 pipe (one DFMA per two clocks and scheduler): the first version
// (8 warps, 8 tokens per warp, F2F conversions, a validity branch per token) spent a third of that
// pipe on the conversions of the query components and about as many issue slots on loop overhead as
// on DFMAs (27 % of the pipe's rate).  Now the conversions run on the integer pipe (f2d_exact) and
// are amortised over 16 document tokens per lane, 128-thread CTAs keep three CTAs per SM.
constexpr int MSX_THREADS = 128;
constexpr int MSX_WARPS = MSX_THREADS / 32;

template <int PW>
__global__ void __launch_bounds__(MSX_THREADS)
ms_exact_kernel(const MsExact a) {
  constexpr int TJ = PW * MSX_WARPS;
  extern __shared__ __align__(16) unsigned char msx_smem[];
  double* s_doc = reinterpret_cast<double*>(msx_smem);                  // [TJ][d]
  double* s_max = s_doc + (size_t)TJ * a.d;                             // [MSX_WARPS][Tq]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f = blockIdx.y;
  const long long q = a.qsel ? a.qsel[f] : f;
  const long long slot = blockIdx.x;
  const long long doc = a.cand ? a.cand[f * a.nc + slot] : slot;
  const double NINF = __longlong_as_double(0xfff0000000000000LL);
  if (doc < 0 || doc >= a.nd) {
    if (threadIdx.x == 0) a.out[f * a.nc + slot] = -NINF;          // padding candidate: +inf key
    return;
  }
  const float* qrow0 = a.qtok + q * a.Tq * a.d;
  const int valid = a.doc_len ? max(0, min(a.doc_len[doc], a.Td)) : a.Td;
  if (a.cand_key) {
    // the band test first (two loads): most candidates of a query end here
    const int kk = (int)min((long long)a.k, a.nc);
    const double E = (double)a.eq[f];
    if ((double)a.cand_key[f * a.nc + slot] > (double)a.cand_key[f * a.nc + kk - 1] + 2.0 * E) {
      if (threadIdx.x == 0) a.out[f * a.nc + slot] = -NINF;
      return;
    }
  }
  constexpr int MAXG = 4;                          // Tq <= 128: up to four tokens per lane
  const int groups = (a.Tq + 31) / 32;
  const int d4 = a.d >> 2;
  double best[MAXG];
#pragma unroll
  for (int g = 0; g < MAXG; ++g) best[g] = NINF;
  for (int tile0 = 0; tile0 < valid; tile0 += TJ) {
    const int nt = min(TJ, valid - tile0);
    __syncthreads();
    // stage the tile: a warp converts one token per trip (float4 reads, 16-byte stores of doubles)
#pragma unroll 1
    for (int j = warp; j < TJ; j += MSX_WARPS) {
      const float4* src = reinterpret_cast<const float4*>(a.docs + (doc * a.Td + tile0 + j) * a.d);
      double* dst = s_doc + (size_t)j * a.d;
      for (int c4 = lane; c4 < d4; c4 += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nt) v = __ldg(src + c4);
        *reinterpret_cast<double2*>(dst + 4 * c4) = make_double2(f2d_exact(v.x), f2d_exact(v.y));
        *reinterpret_cast<double2*>(dst + 4 * c4 + 2) = make_double2(f2d_exact(v.z), f2d_exact(v.w));
      }
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      if (g >= groups) break;
      const int t = g * 32 + lane;
      const float* qrow = qrow0 + (long long)min(t, a.Tq - 1) * a.d;
      double acc[PW];
#pragma unroll
      for (int jj = 0; jj < PW; ++jj) acc[jj] = 0.0;
      const double* dbase = s_doc + (size_t)warp * a.d;
      const int dstep = MSX_WARPS * a.d;           // doubles between two tokens of this warp
      int cb = 0;
      // (the query components of the next pass are fetched a pass ahead: a lane reads its own token's row,
      //  32 bytes per pass, and the L1 round trip was the kernel's first stall reason)
      float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
      if (a.d >= 8) {
        qa = __ldg(reinterpret_cast<const float4*>(qrow));
        qb = __ldg(reinterpret_cast<const float4*>(qrow + 4));
      }
#pragma unroll 1
      for (; cb + 8 <= a.d; cb += 8) {
        float4 na = qa, nb = qb;
        if (cb + 16 <= a.d) {
          na = __ldg(reinterpret_cast<const float4*>(qrow + cb + 8));
          nb = __ldg(reinterpret_cast<const float4*>(qrow + cb + 12));
        }
        const double q0 = f2d_exact(qa.x), q1 = f2d_exact(qa.y), q2 = f2d_exact(qa.z), q3 = f2d_exact(qa.w);
        const double q4 = f2d_exact(qb.x), q5 = f2d_exact(qb.y), q6 = f2d_exact(qb.z), q7 = f2d_exact(qb.w);
        const double2* dv = reinterpret_cast<const double2*>(dbase + cb);
#pragma unroll
        for (int jj = 0; jj < PW; ++jj) {
          const double2 d0 = dv[0], d1 = dv[1], d2 = dv[2], d3 = dv[3];
          double x = acc[jj];
          x = fma(q0, d0.x, x); x = fma(q1, d0.y, x); x = fma(q2, d1.x, x); x = fma(q3, d1.y, x);
          x = fma(q4, d2.x, x); x = fma(q5, d2.y, x); x = fma(q6, d3.x, x); x = fma(q7, d3.y, x);
          acc[jj] = x;
          dv += dstep / 2;
        }
        qa = na; qb = nb;
      }
      if (cb < a.d) {                              // d % 8 == 4: the last four components
        const float4 qa = __ldg(reinterpret_cast<const float4*>(qrow + cb));
        const double q0 = f2d_exact(qa.x), q1 = f2d_exact(qa.y), q2 = f2d_exact(qa.z), q3 = f2d_exact(qa.w);
        const double2* dv = reinterpret_cast<const double2*>(dbase + cb);
#pragma unroll
        for (int jj = 0; jj < PW; ++jj) {
          const double2 d0 = dv[0], d1 = dv[1];
          double x = acc[jj];
          x = fma(q0, d0.x, x); x = fma(q1, d0.y, x); x = fma(q2, d1.x, x); x = fma(q3, d1.y, x);
          acc[jj] = x;
          dv += dstep / 2;
        }
      }
#pragma unroll
      for (int jj = 0; jj < PW; ++jj)
        if (warp + MSX_WARPS * jj < nt) best[g] = fmax(best[g], acc[jj]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < MAXG; ++g) {
    const int t = g * 32 + lane;
    if (g < groups && t < a.Tq) s_max[warp * a.Tq + t] = best[g];
  }
  __syncthreads();
  if (warp == 0) {
    double sum = 0.0;
    for (int t = lane; t < a.Tq; t += 32) {
      double b = s_max[t];
      for (int w = 1; w < MSX_WARPS; ++w) b = fmax(b, s_max[w * a.Tq + t]);
      sum += b;
    }
    sum = warp_sum(sum);
    if (lane == 0) a.out[f * a.nc + slot] = -sum;
  }
}

// E_q of every query (ms_error_bound of sum_t |q_t|), once per call: the band test of ms_exact_kernel and
// the certificate of ms_final_kernel read it.  One warp per query.
__global__ void __launch_bounds__(256)
ms_eq_kernel(const float* __restrict__ qtok, long long nq, int Tq, int d, const Scalars* __restrict__ sc, float c_rel,
             float* __restrict__ eq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * 8 + warp;
  if (q >= nq) return;
  double sn = 0.0;
  for (int t = 0; t < Tq; ++t) {
    float n2 = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = qtok[(q * Tq + t) * d + c]; n2 = fmaf(v, v, n2); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    sn += sqrt((double)n2);
  }
  // (rounded up: the bound must not shrink in the fp32 store)
  if (lane == 0) eq[q] = __double2float_ru(ms_error_bound(sc, c_rel, Tq, d, sn * 1.000001));
}

// per query: order the candidates by (exact score descending, document id), emit k, certify
struct MsFinal {
  const float* qtok; long long nq; int Tq, d, k; long long nc, nd;
  const float* cand_key;      // [nq][nc] approximate -(score), ascending
  const long long* cand_id;   // [nq][nc]
  const double* exact;        // [nq][nc] -(exact score)
  const Scalars* sc;
  float c_rel;
  long long id_offset;
  float* out_score; long long* out_idx;
  int* fail_list; Scalars* scw;
};

__global__ void __launch_bounds__(128)
ms_final_kernel(const MsFinal a) {
  __shared__ double s_key[1024];
  __shared__ long long s_id[1024];
  __shared__ float s_red[4];
  const long long q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // E_q = sum over the tokens of the dot-product bound (key_error_bound's direct-dot terms)
  float sn = 0.f;
  for (int t = warp; t < a.Tq; t += 4) {
    float n2 = 0.f;
    for (int c = lane; c < a.d; c += 32) { const float v = a.qtok[(q * a.Tq + t) * a.d + c]; n2 = fmaf(v, v, n2); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    sn += sqrtf(n2);
  }
  if (lane == 0) s_red[warp] = sn;
  const int n2p = max(next_pow2((int)a.nc), 2);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  for (int i = threadIdx.x; i < n2p; i += blockDim.x) {
    double kv = INF; long long id = 0x7fffffffffffffffLL;
    if (i < a.nc) {
      const long long cid = a.cand_id[q * a.nc + i];
      if (cid >= 0) {
        kv = a.exact[q * a.nc + i];
        if (kv != kv) kv = INF;
        if (kv < INF) id = cid;                      // (+inf: outside the band, not scored; empty documents)
      }
    }
    s_key[i] = kv; s_id[i] = id;
  }
  __syncthreads();
  const double sum_norm = (double)s_red[0] + s_red[1] + s_red[2] + s_red[3];
  const double E = ms_error_bound(a.sc, a.c_rel, a.Tq, a.d, sum_norm * 1.000001);
  // monitor: the approximate scores must lie within E of the exact ones
  float ratio = 0.f;
  for (int i = threadIdx.x; i < a.nc; i += blockDim.x) {
    const double ex = a.exact[q * a.nc + i];
    const long long cid = a.cand_id[q * a.nc + i];
    if (cid >= 0 && ex == ex && ex < INF && ex > -INF)
      ratio = fmaxf(ratio, (float)(fabs((double)a.cand_key[q * a.nc + i] - ex) / E));
  }
  block_bitonic_sort(s_key, s_id, n2p);
  for (int j = threadIdx.x; j < a.k; j += blockDim.x) {
    const bool pad = j >= n2p || s_id[j] == 0x7fffffffffffffffLL;
    a.out_score[q * a.k + j] = pad ? -__int_as_float(0x7f800000) : (float)(-s_key[j]);
    a.out_idx[q * a.k + j] = pad ? -1 : s_id[j] + a.id_offset;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ratio = fmaxf(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
  if (lane == 0) atomicMax(&a.scw->err_ratio, __float_as_uint(ratio));
  if (threadIdx.x == 0) {
    // certificate: the candidates are the nc best approximate scores; a document outside has an
    // approximate key >= the last candidate's, so it cannot matter if that lies above a_k + 2 E
    bool certified = a.nc >= a.nd;
    if (!certified && a.nc > a.k) {
      const double a_k = (double)a.cand_key[q * a.nc + a.k - 1];
      const double a_last = (double)a.cand_key[q * a.nc + a.nc - 1];
      certified = a_last > a_k + 2.0 * E;
    }
    if (!certified) { const int slot = atomicAdd(&a.scw->fail_cnt, 1); a.fail_list[slot] = (int)q; }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
size_t topk_from_keys_workspace_bytes(long long n, long long F, int k);
int topk_from_keys(const double* keys, long long n, long long stride, long long F, int k, const int* qsel,
                   long long id_offset, float* out_key, long long* out_idx, void* ws, size_t ws_bytes, cudaStream_t s);

__global__ void ms_negate_rows_kernel(float* out, const int* rows, int nrows, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows * k) { float* p = out + (long long)rows[i / k] * k + (i % k); *p = -*p; }
}

struct MsPlan { int dh, pitch, nKB, kLast, N, nQT, G, nUnits; long long nc; long long fb_rows; };

static MsPlan ms_plan(long long nq, int Tq, long long nd, int Td, int d, int k) {
  MsPlan pl{};
  pl.dh = (d + 7) / 8 * 8;
  pl.pitch = (pl.dh + 63) / 64 * 64;
  pl.nKB = (pl.dh + 63) / 64;
  pl.kLast = (pl.dh - (pl.nKB - 1) * 64 + 15) / 16;
  pl.N = (Td + 31) / 32 * 32;
  pl.nQT = (int)((nq * Tq + TQ - 1) / TQ);
  pl.G = 16;
  const long long chunks = (nd + pl.G - 1) / pl.G;
  pl.nUnits = (int)(chunks * pl.nQT);
  long long nc = 4ll * k;
  if (nc < 256) nc = 256;
  if (nc > 1024) nc = 1024;
  if (nc > nd) nc = nd;
  pl.nc = nc;
  // uncertified queries are recomputed against all documents, fb_rows of them at a time
  pl.fb_rows = (256ll << 20) / (nd * 8) + 1;
  if (pl.fb_rows > 64) pl.fb_rows = 64;
  return pl;
}

bool maxsim_supported(long long nq, int Tq, long long nd, int Td, int d, int k) {
  return nq >= 1 && nd >= 1 && Tq >= 1 && Tq <= 128 && (Tq & (Tq - 1)) == 0 && Td >= 1 && Td <= TB && d >= 4 && d % 4 == 0 &&
         d <= 1024 && k >= 1 && k <= 256 && k <= nd && nd * ((Td + 31) / 32 * 32) < (1ll << 31) && nq * Tq < (1ll << 31);
}

size_t maxsim_workspace_bytes(long long nq, int Tq, long long nd, int Td, int d, int k) {
  const MsPlan pl = ms_plan(nq, Tq, nd, Td, d, k);
  size_t b = 0;
  b += align_up((size_t)nq * Tq * pl.pitch * 2, 256);
  b += align_up((size_t)nd * pl.N * pl.pitch * 2, 256);
  b += align_up((size_t)nq * nd * 8, 256);
  b += align_up((size_t)nq * pl.nc * 4, 256) + align_up((size_t)nq * pl.nc * 8, 256) * 2;
  b += align_up((size_t)nq * 4, 256) * 2 + 4096;
  b += align_up((size_t)pl.fb_rows * nd * 8, 256);
  const size_t t1 = topk_from_keys_workspace_bytes(nd, nq, (int)pl.nc), t2 = topk_from_keys_workspace_bytes(nd, pl.fb_rows, k);
  b += (t1 > t2 ? t1 : t2) + 4096;
  return b;
}

int maxsim_topk(const float* qtok, long long nq, int Tq, const float* docs, long long nd, int Td, const int* doc_len, int d,
                int k, long long id_offset, float* out_score, long long* out_idx, void* ws, size_t ws_bytes,
                nwk_knn_stats* stats, cudaStream_t s) {
  if (!maxsim_supported(nq, Tq, nd, Td, d, k))
    return fail(NWK_ERR_ARG, "maxsim: unsupported shape (Tq a power of two <= 128, Td <= 256, d %% 4 == 0, k <= min(256, docs))");
  if (ws_bytes < maxsim_workspace_bytes(nq, Tq, nd, Td, d, k)) return fail(NWK_ERR_WORKSPACE, "maxsim workspace too small");
  const int sms = sm_count();
  const MsPlan pl = ms_plan(nq, Tq, nd, Td, d, k);
  Arena ar(ws, ws_bytes);
  __half* qh = ar.take<__half>((size_t)nq * Tq * pl.pitch);
  __half* dhh = ar.take<__half>((size_t)nd * pl.N * pl.pitch);
  double* neg = ar.take<double>((size_t)nq * nd);
  float* cand_key = ar.take<float>((size_t)nq * pl.nc);
  long long* cand_id = ar.take<long long>((size_t)nq * pl.nc);
  double* exact = ar.take<double>((size_t)nq * pl.nc);
  int* fail_list = ar.take<int>((size_t)nq);
  float* eq = ar.take<float>((size_t)nq);
  Scalars* sc = ar.take<Scalars>(1);
  double* fb = ar.take<double>((size_t)pl.fb_rows * nd);
  ar.off = align_up(ar.off, 256);
  void* tws = ar.base + ar.off;
  const size_t tws_bytes = ws_bytes - ar.off;
  const long long launches0 = g_launches.load();

  cudaEvent_t ev[6];
  for (int i = 0; i < 6; ++i) cudaEventCreate(&ev[i]);
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 6; ++i) cudaEventDestroy(e[i]); } } ev_guard{ev};
  cudaEventRecord(ev[0], s);
  NWK_CUDA(cudaMemsetAsync(sc, 0, sizeof(Scalars), s));
  qmax_kernel<<<sms, 256, 0, s>>>(qtok, nq * Tq * d, &sc->qabsmax);
  NWK_LAUNCH_CHECK();
  qmax_kernel<<<sms * 8, 256, 0, s>>>(docs, nd * Td * d, &sc->babsmax);
  NWK_LAUNCH_CHECK();
  qprep_kernel<<<sms * 2, 256, 0, s>>>(qtok, nq * Tq, d, pl.dh, pl.pitch, NWK_METRIC_DOT, 0, sc, qh);
  NWK_LAUNCH_CHECK();
  {
    long long blocks = (nd * pl.N + 7) / 8;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    ms_docprep_kernel<<<(unsigned)blocks, 256, 0, s>>>(docs, nd, Td, pl.N, d, pl.pitch, sc, dhh);
    NWK_LAUNCH_CHECK();
  }
  cudaEventRecord(ev[1], s);
  CUtensorMap tm_q, tm_d;
  NWK_TRY(make_tmap(&tm_q, qh, nq * Tq, pl.dh, true, TQ, pl.pitch));
  NWK_TRY(make_tmap(&tm_d, dhh, nd * pl.N, pl.dh, true, pl.N, pl.pitch));
  if (Tq > 32) NWK_CUDA(cudaMemsetAsync(neg, 0, (size_t)nq * nd * 8, s));
  MsParams p{};
  p.nKB = pl.nKB; p.kLast = pl.kLast; p.nq = nq; p.nd = nd; p.Tq = Tq; p.Td = Td; p.N = pl.N; p.doc_len = doc_len;
  p.nQT = pl.nQT; p.G = pl.G; p.nUnits = pl.nUnits; p.sc = sc; p.neg_score = neg; p.err = &sc->err;
  {
    const int smem = STAGES * STAGE_BYTES + 256 + 1024;
    if (!dev_state()->ms_optin) {
      NWK_CUDA(cudaFuncSetAttribute(maxsim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      dev_state()->ms_optin = 1;
    }
    maxsim_kernel<<<pl.nUnits < sms ? pl.nUnits : sms, MS_THREADS, smem, s>>>(tm_q, tm_d, p);
    NWK_LAUNCH_CHECK();
  }
  cudaEventRecord(ev[2], s);
  NWK_TRY(topk_from_keys(neg, nd, nd, nq, (int)pl.nc, nullptr, 0, cand_key, cand_id, tws, tws_bytes, s));
  cudaEventRecord(ev[3], s);

  // document tile of the exact kernel: TJ = 4 PW tokens as doubles, at most 64 KB (three CTAs per SM)
  int pw = 16;
  while (pw > 2 && (size_t)pw * MSX_WARPS * d * 8 > 64 * 1024) pw >>= 1;
  if (const char* e = getenv("NWK_MSX_PW")) {                      // experiment switch: tokens per lane (16, 8, 4, 2)
    const int want = atoi(e);
    if ((want == 16 || want == 8 || want == 4 || want == 2) && want <= pw) pw = want;
  }
  const int ex_smem = pw * MSX_WARPS * d * 8 + MSX_WARPS * Tq * 8 + 64;
  if (ex_smem > dev_state()->ms_ex_optin) {
    NWK_CUDA(cudaFuncSetAttribute(ms_exact_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ex_smem));
    NWK_CUDA(cudaFuncSetAttribute(ms_exact_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ex_smem));
    NWK_CUDA(cudaFuncSetAttribute(ms_exact_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ex_smem));
    NWK_CUDA(cudaFuncSetAttribute(ms_exact_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ex_smem));
    dev_state()->ms_ex_optin = ex_smem;
  }
  auto launch_exact = [&](const MsExact& es, dim3 grid) -> int {
    if (pw == 16) ms_exact_kernel<16><<<grid, MSX_THREADS, ex_smem, s>>>(es);
    else if (pw == 8) ms_exact_kernel<8><<<grid, MSX_THREADS, ex_smem, s>>>(es);
    else if (pw == 4) ms_exact_kernel<4><<<grid, MSX_THREADS, ex_smem, s>>>(es);
    else ms_exact_kernel<2><<<grid, MSX_THREADS, ex_smem, s>>>(es);
    NWK_LAUNCH_CHECK();
    return NWK_OK;
  };
  const float c_rel = (float)(1.0 / 1024.0 * 1.001 + ((double)d / 16.0 + 2.0) / 4194304.0);
  ms_eq_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, s>>>(qtok, nq, Tq, d, sc, c_rel, eq);
  NWK_LAUNCH_CHECK();
  MsExact e{};
  e.qtok = qtok; e.docs = docs; e.doc_len = doc_len; e.nq = nq; e.nd = nd; e.Tq = Tq; e.Td = Td; e.d = d;
  for (long long q0 = 0; q0 < nq; q0 += 32768) {
    const long long nf = std::min<long long>(32768, nq - q0);
    MsExact es = e;
    es.qtok = qtok + q0 * Tq * d; es.cand = cand_id + q0 * pl.nc; es.nc = pl.nc; es.out = exact + q0 * pl.nc;
    es.cand_key = cand_key + q0 * pl.nc; es.k = k; es.eq = eq + q0;
    NWK_TRY(launch_exact(es, dim3((unsigned)pl.nc, (unsigned)nf)));
  }
  MsFinal f{};
  f.qtok = qtok; f.nq = nq; f.Tq = Tq; f.d = d; f.k = k; f.nc = pl.nc; f.nd = nd; f.cand_key = cand_key; f.cand_id = cand_id;
  f.exact = exact; f.sc = sc; f.scw = sc; f.id_offset = id_offset; f.out_score = out_score; f.out_idx = out_idx;
  f.fail_list = fail_list;
  f.c_rel = c_rel;
  ms_final_kernel<<<(unsigned)nq, 128, 0, s>>>(f);
  NWK_LAUNCH_CHECK();
  cudaEventRecord(ev[4], s);

  Scalars h{};
  NWK_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scalars), cudaMemcpyDeviceToHost, s));
  NWK_CUDA(cudaStreamSynchronize(s));
  if (h.err != 0) return fail(NWK_ERR_INTERNAL, "tcgen05 pipeline watchdog expired (role code %d)", h.err);
  float err_ratio;
  memcpy(&err_ratio, &h.err_ratio, 4);
  if (err_ratio > 1.0f) return fail(NWK_ERR_INTERNAL, "maxsim scores exceeded the proven error bound (ratio %.3f)", err_ratio);
  // uncertified queries: exact scores against every document
  for (long long done = 0; done < h.fail_cnt; done += pl.fb_rows) {
    const long long nf = std::min<long long>(pl.fb_rows, h.fail_cnt - done);
    MsExact es = e;
    es.qsel = fail_list + done; es.cand = nullptr; es.nc = nd; es.out = fb;
    NWK_TRY(launch_exact(es, dim3((unsigned)nd, (unsigned)nf)));
    NWK_TRY(topk_from_keys(fb, nd, nd, nf, k, fail_list + done, id_offset, out_score, out_idx, tws, tws_bytes, s));
    ms_negate_rows_kernel<<<(unsigned)((nf * k + 255) / 256), 256, 0, s>>>(out_score, fail_list + done, (int)nf, k);
    NWK_LAUNCH_CHECK();
  }
  cudaEventRecord(ev[5], s);
  NWK_CUDA(cudaStreamSynchronize(s));
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    // stages of this path: prep = fp16 copies, filter = the tcgen05 GEMM with the fused max / sum,
    // sample = top-4k of the approximate scores, finish = exact re-scoring + ordering, fallback
    cudaEventElapsedTime(&stats->ms_prep, ev[0], ev[1]);
    cudaEventElapsedTime(&stats->ms_filter, ev[1], ev[2]);
    cudaEventElapsedTime(&stats->ms_sample, ev[2], ev[3]);
    cudaEventElapsedTime(&stats->ms_finish, ev[3], ev[4]);
    cudaEventElapsedTime(&stats->ms_fallback, ev[4], ev[5]);
    cudaEventElapsedTime(&stats->ms_total, ev[0], ev[5]);
    stats->path = 3;
    stats->n_fallback_queries = h.fail_cnt;
    stats->max_candidates = (int)pl.nc;
    stats->gpu_launches = (int)(g_launches.load() - launches0);
    stats->max_err_ratio = err_ratio;
  }
  return NWK_OK;
}
