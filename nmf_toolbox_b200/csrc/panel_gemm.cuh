// panel_gemm: the one tensor-core kernel behind every large contraction of the
// NMF multiplicative updates (SURVEY.md section 7, K2/K4 and the Gram GEMMs).
//
//   acc0[128 x bn] = sum over k-blocks of  X0[r0:r0+128, kb] * Y0[n0:n0+bn, kb]^T   (phase 0)
//   acc1[128 x bn] = sum over k-blocks of  X1[r0:r0+128, kb] * Y1[n0:n0+bn, kb]^T   (phase 1)
//
// X is the "panel" operand (V, W or H seen so that its rows are the output
// rows), Y the small factor (H, W or a K x K Gram matrix).  Operands arrive by
// TMA (128-byte swizzle) through a 4-stage mbarrier ring and tcgen05.mma
// kind::tf32 accumulates into TMEM.
//
// Two-level accumulation.  The tensor core truncates (does not round) when it
// adds into the fp32 accumulator, which gives a systematic relative bias of
// about 5e-8 per accumulation step (measured on B200: 1.1e-4 after the 2048
// steps of a 16384-long contraction).  Phase 0 is therefore cut into chunks of
// kChunkKb k-blocks (64 MMA steps) that ping-pong between the two 256-column
// TMEM buffers; eight epilogue warps drain each finished chunk with tcgen05.ld
// and add it, round-to-nearest, into fp32 registers while the next chunk is
// already accumulating.  Phase 1 (short: the K x K Gram products) is one chunk.
//
// Epilogues (thread = one output row, accumulator columns in registers):
//   EPI_STORE    raw sums (optionally one slab per split-K slice)
//   EPI_HUPDATE  fused multiplicative H update, H <- H .* (N ./ max(D + lambda, eps))
//                (nmf.m:180-181,199): D is the second accumulator (Euclidean,
//                D = (W'W) H) or a per-basis vector (KL, D = W' * ones = column
//                sums of W, nmf.m:183-184); N and D never exist in HBM.
//   EPI_RECON    acc is a tile of V_hat = W*H (ReconstructFromDecomposition.m:31): store it
//   EPI_RESID    sum of (V - V_hat).^2 over the tile (nmf.m:208, nmfsc.m:161)
//   EPI_KLQ      Q = V ./ V_hat written tf32-rounded (nmf.m:152,183) and, on request,
//                sum(V .* log(V_hat)), sum(V_hat) for the KL cost (nmf.m:210)
//   EPI_ABQ      the two element-wise weight matrices of the IS / AB updates (nmf.m:154-164,
//                185-195), tf32-rounded: Qn multiplies H' / W' in the negative gradient, Qp in
//                the positive one; on request the sum of the per-element divergence (nmf.m:211-214)
// Every kernel of an iteration loop starts by reading *stop: once the cost
// kernel has detected convergence (nmf.m:221-224) the launches already queued
// behind it become no-ops, so the host never has to synchronise per iteration.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2..9 = epilogue (TMEM lane quarter = warp & 3, column half =
// (warp - 2) >> 2).
#pragma once
#include "ptx.cuh"

namespace nmfb {

constexpr int kTileM = 128;   // output rows per CTA = TMEM lanes
constexpr int kBlockK = 32;   // fp32 elements per 128-byte swizzle row
constexpr int kUmmaK = 8;     // tf32 MMA K
constexpr int kStages = 4;
#ifndef NMFB_PAIR_STAGES
#define NMFB_PAIR_STAGES 6  // operand ring of the CTA-pair kernels: 6 x 32 KB (7 fits the 227 KB limit as well)
#endif
constexpr int kMaxN = 256;    // widest accumulator (fp32 TMEM columns)
constexpr int kChunkKb = 16;  // k-blocks accumulated in TMEM before promotion to registers (64 MMA steps:
                              // accumulate-truncation bias ~3e-6 relative, 2.5 % faster than 8)
constexpr int kStageBytesX = kTileM * 128;
constexpr int kStageBytesY = kMaxN * 128;
constexpr int kStageBytes = kStageBytesX + kStageBytesY;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kGemmSmemBytes = kStages * kStageBytes + 1024;

// CTA-pair variant (CG = 2, tcgen05 cta_group::2): two CTAs of a cluster on neighbouring SMs
// compute one 256 x N tile.  Each CTA loads its own 128 rows of X but only HALF of the Y
// slab (the tensor core reads the other half from the partner's shared memory), so a stage
// is 32 KB instead of 48 KB - a third less L2->SM traffic per flop and room for 6 stages.
// Cluster-of-four variant (CG = 4): two CTA pairs on neighbouring row tiles read the SAME Y slab, so each
// CTA fetches only a QUARTER of it and TMA-multicasts that quarter to the CTA holding the same half in the
// other pair.  Per CTA and stage the L2 -> SM traffic drops from 16 + 16 KB to 16 + 8 KB; the shared-memory
// layout (and everything downstream of the full barriers) is the pair kernel's.  A stage is reusable only
// when BOTH pairs have retired its MMAs (the multicast writes into the other pair's buffers too), so the
// empty barriers collect one tcgen05.commit from each pair leader.
template <int CG>
struct TileCfg {
  static constexpr int pair = CG >= 2 ? 2 : 1;  // CTAs that share one MMA (cta_group)
  static constexpr int stages = CG >= 2 ? NMFB_PAIR_STAGES : kStages;
  static constexpr int ybytes = kStageBytesY / pair;
  static constexpr int stage_bytes = kStageBytesX + ybytes;
  static constexpr int smem_bytes = stages * stage_bytes + 1024;
};
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxGroups = kMaxN / 16 / 2;  // 16-column groups per epilogue thread

// MATLAB's eps (2^-52) - the reference clamps denominators with the double
// eps even for single data (nmf.m:168,199); representable in fp32.
#define NMFB_EPS 2.220446049250313e-16f

enum { EPI_STORE = 0, EPI_HUPDATE = 1, EPI_RECON = 2, EPI_RESID = 3, EPI_KLQ = 4, EPI_ABQ = 5 };
enum { ABQ_IS = 0, ABQ_AB = 1, ABQ_AB_DUAL = 2 };

struct GemmArgs {
  int rows;          // valid output rows (rows of X)
  int ncols;         // output columns (multiple of 32; rows of Y beyond its extent read as zero)
  int ncols_valid;   // EPI_RECON / EPI_RESID / EPI_KLQ: columns that exist in V (<= ncols)
  int box_n;         // rows of the Y TMA box = min(ncols, tile width)
  int tile_n;        // output columns per CTA (0 = 256): narrower tiles put more CTAs on a problem with few
                     // row tiles while keeping fused epilogues (every CTA still runs the whole contraction)
  int nkb0;          // phase-0 k-blocks in total (over all splits and segments)
  int nkb_seg;       // phase 0 may be a sum over up to 3 operand pairs ("segments", e.g. the
                     // hi/lo terms of a split-tf32 product); k-blocks per segment
  int chunk_kb;      // k-blocks accumulated in TMEM before promotion to registers (<= 0: default)
  int kb_per_split;  // phase-0 k-blocks per split (blockIdx.z)
  int nkb1;          // phase-1 k-blocks (split 0 only); 0 = no second accumulator
  int xmn0, xmn1;    // X operand of phase 0 / 1 is MN-major (rows contiguous) instead of K-major
  int ymn0, ymn1;    // same for the Y operand (its rows = output columns are contiguous)
  // EPI_STORE: out[z * split_stride + col * ldo + row]
  float* out0;
  float* out1;
  long long ldo;
  long long split_stride;
  // EPI_HUPDATE: H master (fp32) and its tf32-rounded operand copies
  float* Hm;     // [ncols][ldh]  row = basis index k, contiguous along samples
  float* Hr32;   // same layout, tf32-rounded
  float* Hc32;   // [rows][ldc]   contiguous along k (optional, may be null)
  long long ldh;
  long long ldc;
  float lambda;
  const float* dvec;  // EPI_HUPDATE: if non-null, D[col] = dvec[col] replaces the second accumulator
  double* scal;       // EPI_HUPDATE: scal[0] += sum(N .* tf32(Hnew)), scal[1] += sum(Hnew)
                      // EPI_RESID:   scal[0] += sum((V - acc)^2)
                      // EPI_KLQ:     scal[0] += sum(V .* log(acc)), scal[1] += sum(acc)   (if want_cost)
                      // EPI_ABQ:     scal[0] += sum of the per-element divergence          (if want_cost)
  // EPI_RECON / EPI_RESID / EPI_KLQ: element (row, col) of the m x n problem lives at [col * ldv + row]
  const float* Vsrc;
  float* Qout;        // EPI_KLQ: Q; EPI_RECON: V_hat; EPI_ABQ: Qn
  float* Qout2;       // EPI_ABQ: Qp
  int ab_mode;        // EPI_ABQ: ABQ_IS / ABQ_AB / ABQ_AB_DUAL (alpha == 0, nmf.m:124-128)
  int ab_cost_kl;     // EPI_ABQ: the summed divergence is KL's (cnmf.m:243 reaches KL as alpha = 1, beta = 0)
  float ab_alpha, ab_beta;
  long long ldv;
  int want_cost;
  int freeze;         // EPI_HUPDATE: leave H untouched (H_fixed, nmf.m:177) but still form the sums
  int h_prefetch;     // EPI_HUPDATE: the producer stages the tile of the H master in shared memory by
                      // TMA (map tmH) as soon as the operand ring is idle, instead of the epilogue
                      // fetching it with a chain of dependent global loads after the last MMA
  const int* stop;    // device flag: non-zero = iteration loop already converged, do nothing
  // Phase 1 (the short second accumulator) may depend on a small kernel that runs CONCURRENTLY
  // on another stream (a Gram matrix): the producer waits for *gate >= gate_value right before
  // its first phase-1 load.  Only valid when the grid leaves SMs free for that kernel.
  const unsigned int* gate;
  unsigned int gate_value;
  // Tail helpers (CTA-pair kernels, one column chunk, no split-K): a grid of `sk_tiles` row tiles that
  // leaves SMs idle (64 pair tiles on 148 SMs) is extended by `sk_helpers` CTA pairs.  The pair of row
  // tile t (the "primary") contracts k-blocks [0, sk_kp) only; helper pair h contracts the tail
  // [sk_kp, nkb0) of the tiles h, h + sk_helpers, ... one after the other, leaves each partial sum in
  // sk_part ([tile][column][256 rows], L2-resident) and publishes sk_flag[2 * tile + cta] = sk_epoch.  The
  // primary adds the partial to its own sum (registers) before its epilogue, so the fused epilogues are
  // unchanged and every SM works through the whole launch (stream-K with a fixed split).
  int sk_helpers;
  int sk_tiles;
  int sk_kp;
  float* sk_part;
  unsigned int* sk_flag;
  unsigned int sk_epoch;
};

// EPI_ABQ: 16 columns of one row.  The mode is uniform over the launch, so the three variants are
// separate straight-line blocks (only one of them is ever fetched).
template <int MODE>
__device__ __forceinline__ void abq_group_mode(const GemmArgs& a, const float (&v)[16], const float* sum, int ncols,
                                               float* qn_out, float* qp_out, float& s0) {
  const float al = a.ab_alpha, be = a.ab_beta;
  // exponents of log2 V and log2 V_hat: Qn = V^c0 V_hat^c1, Qp = V^c2 V_hat^c3
  const float c0 = MODE == ABQ_AB ? al : al - 1.f, c1 = MODE == ABQ_AB ? be - 1.f : be;
  const float c2 = MODE == ABQ_AB ? 0.f : al + be - 1.f, c3 = MODE == ABQ_AB ? al + be - 1.f : 0.f;
  const float rs = 1.f / (al + be);  // alpha + beta == 0: Inf, as the reference's division
  const long long ldv = a.ldv;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (t < ncols) {
      const float sv = sum[t], vv = v[t];
      float qn, qp;
      if (MODE == ABQ_IS) {  // nmf.m:155-156,186-187
        const float r = fast_rcp(sv);
        qp = r;
        qn = vv * r * r;
        if (a.want_cost) s0 += 0.6931471805599453f * fast_lg2(sv * fast_rcp(vv)) + vv * r - 1.f;  // nmf.m:212
      } else {               // nmf.m:159-163,190-194: powers as exp2(c log2 x); x.^0 == 1 (no 0 * inf)
        const float lv = fast_lg2(vv), ls = fast_lg2(sv);
        auto term = [](float c, float l) { return c == 0.f ? 0.f : c * l; };
        qn = fast_ex2(term(c0, lv) + term(c1, ls));
        qp = fast_ex2(term(c2, lv) + term(c3, ls));
        if (a.want_cost) {
          if (a.ab_cost_kl)  // cnmf.m:243
            s0 += vv * 0.6931471805599453f * (lv - ls) - vv + sv;
          else               // nmf.m:214 (the prefactor is applied by the cost kernel)
            s0 += fast_ex2(term(al, lv) + term(be, ls)) -
                  (al * fast_ex2(term(al + be, lv)) + be * fast_ex2(term(al + be, ls)) + be) * rs;
        }
      }
      qn_out[t * ldv] = tf32_rn(qn);
      qp_out[t * ldv] = tf32_rn(qp);
    }
  }
}
template <int MODE, int NG>
__device__ __forceinline__ void abq_tile(const GemmArgs& a, const float (&sum)[NG * 16], int g_count, int col0, int row,
                                         int cols_ok, bool staged, const float* vsm, int vsm_stride, float& s0) {
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g < g_count) {
      const long long off = static_cast<long long>(col0 + g * 16) * a.ldv + row;
      float v[16];
#pragma unroll
      for (int t = 0; t < 16; ++t)
        v[t] = (g * 16 + t < cols_ok) ? (staged ? vsm[(g * 16 + t) * vsm_stride] : __ldg(a.Vsrc + off + t * a.ldv)) : 0.f;
      abq_group_mode<MODE>(a, v, &sum[g * 16], cols_ok - g * 16, a.Qout + off, a.Qout2 + off, s0);
    }
  }
}

// Rolled variant: accumulator groups come from TMEM one at a time (see `direct` in the kernel).
template <int MODE>  // MODE < 0: EPI_KLQ
__device__ __noinline__ void q_tile_direct(const GemmArgs& a, uint32_t taddr, int g_count, int col0, int row,
                                           bool row_ok, int cols_ok, bool staged, const float* vsm, int vsm_stride, float& s0,
                                           float& s1) {
#pragma unroll 1
  for (int g = 0; g < g_count; ++g) {
    float sv[16], v[16];
    tmem_ld16(taddr + g * 16, sv);  // warp-collective: every lane gets here, rows beyond m included
    const long long off = static_cast<long long>(col0 + g * 16) * a.ldv + row;
    const int nc = row_ok ? cols_ok - g * 16 : 0;
#pragma unroll
    for (int t = 0; t < 16; ++t)
      v[t] = (t < nc) ? (staged ? vsm[(g * 16 + t) * vsm_stride] : __ldg(a.Vsrc + off + t * a.ldv)) : 0.f;
    tmem_ld_wait();
    if constexpr (MODE < 0) {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (t < nc) {
          a.Qout[off + t * a.ldv] = tf32_rn(v[t] * fast_rcp(sv[t]));
          if (a.want_cost) {
            s0 = fmaf(v[t], fast_lg2(sv[t]), s0);
            s1 += sv[t];
          }
        }
      }
    } else {
      abq_group_mode<MODE>(a, v, sv, nc, a.Qout + off, a.Qout2 + off, s0);
    }
  }
}

template <int EPI, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
panel_gemm_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmY0,
                  const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmY1,
                  const __grid_constant__ CUtensorMap tmXb, const __grid_constant__ CUtensorMap tmYb,
                  const __grid_constant__ CUtensorMap tmXc, const __grid_constant__ CUtensorMap tmYc,
                  const __grid_constant__ CUtensorMap tmH, const GemmArgs a) {
  using TC = TileCfg<CG>;
  constexpr int kNStages = TC::stages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kNStages];
  __shared__ uint64_t empty_bar[kNStages];
  __shared__ uint64_t tfull_bar[2];   // TMEM buffer holds a finished chunk
  __shared__ uint64_t tempty_bar[2];  // TMEM buffer has been drained
  __shared__ uint64_t h_bar;          // H master tile has landed (EPI_HUPDATE with h_prefetch)
  __shared__ uint32_t tmem_slot;
  __shared__ double red[kEpiWarps][2];

  if (a.stop != nullptr && *a.stop != 0) return;  // uniform across the grid
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;

  constexpr bool kPair = CG >= 2;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;  // rank in the cluster (CG = 4: 0..3, two pairs)
  const uint32_t rank = crank & 1u;                        // position in the CTA pair; 0 = leader (issues the MMAs)
  const uint32_t lead = crank & ~1u;                       // cluster rank of this pair's leader
  // tail helpers (see GemmArgs): work items of this CTA pair
  const bool sk = kPair && CG == 2 && a.sk_helpers > 0;
  const int pair_idx = static_cast<int>(blockIdx.x >> 1);
  const bool helper = sk && pair_idx >= a.sk_tiles;
  const int first_tile = helper ? pair_idx - a.sk_tiles : pair_idx;
  const int n_items = helper ? (a.sk_tiles - first_tile + a.sk_helpers - 1) / a.sk_helpers : 1;
  const int item_rows = sk ? a.sk_helpers * 2 * kTileM : 0;  // row distance between a helper's tiles
  const int r0_first = kPair ? first_tile * (2 * kTileM) + static_cast<int>(rank) * kTileM
                             : static_cast<int>(blockIdx.x) * kTileM;
  const int r0 = r0_first;  // primaries and every kernel without helpers: the one row tile of this CTA
  const int tile_n = a.tile_n > 0 ? a.tile_n : kMaxN;
  const int n0 = blockIdx.y * tile_n;
  const int bn = min(a.ncols - n0, tile_n);  // multiple of 32
  const int split = blockIdx.z;
  const int kb_begin = sk ? (helper ? a.sk_kp : 0) : split * a.kb_per_split;
  const int n0kb = sk ? (helper ? a.nkb0 - a.sk_kp : a.sk_kp)
                      : max(0, min(a.nkb0, kb_begin + a.kb_per_split) - kb_begin);
  const int n1kb = (split == 0 && !helper) ? a.nkb1 : 0;
  const int chunk_kb = a.chunk_kb > 0 ? a.chunk_kb : kChunkKb;
  const int nchunk0 = (n0kb + chunk_kb - 1) / chunk_kb;
  const int nchunks = nchunk0 + (n1kb > 0 ? 1 : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CG == 4 ? 2 : 1);  // CG = 4: both pairs must have retired the stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kEpiWarps * TC::pair);  // the leader's barrier collects both CTAs' epilogues
    }
    mbar_init(&h_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmX0);
    prefetch_tmap(&tmY0);
    if (a.nkb1 > 0) {
      prefetch_tmap(&tmX1);
      prefetch_tmap(&tmY1);
    }
  }
  if (warp == 1) {
    if constexpr (kPair) {
      tmem_alloc_pair(&tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ TMA producer
    // bytes that land per stage in BOTH CTAs of a pair are counted on the leader's barrier
    const int ybox = a.box_n / TC::pair;  // rows of Y in this CTA's shared memory (its half of the slab)
    const uint32_t tx_bytes = TC::pair * (kStageBytesX + ybox * 128);
    const int yrow0 = n0 + static_cast<int>(rank) * (bn / TC::pair);
    // CG = 4: this CTA fetches quarter `yq` of the slab (half of its half) and multicasts it to the CTA with the
    // same position in the other pair
    const int yq = CG == 4 ? static_cast<int>(crank >> 1) : 0;
    const uint16_t ymask = static_cast<uint16_t>(0x5u << rank);
    const int total = n0kb + n1kb;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = 0; item < n_items; ++item) {
      const int r0 = r0_first + item * item_rows;  // a helper pair walks over its row tiles
      for (int it = 0; it < total; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        const uint32_t fb = kPair ? map_to_cta(smem_u32(&full_bar[stage]), lead) : 0u;
        auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint64_t pol) {
          if constexpr (kPair) tma_load_2d_pair(dst, tm, fb, c0, c1, pol);
          else tma_load_2d(dst, tm, &full_bar[stage], c0, c1, pol);
        };
        const bool ph1 = it >= n0kb;
        if (ph1 && it == n0kb && a.gate != nullptr) gate_wait(a.gate, a.gate_value);
        int kb = ph1 ? (it - n0kb) : (kb_begin + it);
        const CUtensorMap* mx = ph1 ? &tmX1 : &tmX0;
        const CUtensorMap* my = ph1 ? &tmY1 : &tmY0;
        if (!ph1 && kb >= a.nkb_seg) {  // second / third operand pair of a segmented phase 0
          const int seg = kb / a.nkb_seg;
          kb -= seg * a.nkb_seg;
          mx = seg == 1 ? &tmXb : &tmXc;
          my = seg == 1 ? &tmYb : &tmYc;
        }
        const uint32_t xs = sbase + stage * TC::stage_bytes;
        const uint32_t ys = xs + kStageBytesX;
        // X is streamed once (evict-first); Y is re-read by every CTA (evict-last).
        if (ph1 ? a.xmn1 : a.xmn0) {
          // rows are the contiguous dimension: four 32(rows) x 32(k) boxes
  #pragma unroll
          for (int q = 0; q < 4; ++q) load(xs + q * 4096, mx, r0 + q * 32, kb * kBlockK, kEvictFirst);
        } else {
          load(xs, mx, kb * kBlockK, r0, kEvictFirst);
        }
        if constexpr (CG == 4) {
          if (ph1 ? a.ymn1 : a.ymn0) {  // 32-row boxes: this CTA's half of the boxes of its half
            const int nq = ybox >> 6;
            for (int q = yq * nq; q < (yq + 1) * nq; ++q)
              tma_load_2d_pair_mc(ys + q * 4096, my, &full_bar[stage], ymask, yrow0 + q * 32, kb * kBlockK, kEvictLast);
          } else {  // one box of ybox / 2 rows (the tensor map was built with box_n / 4 rows)
            tma_load_2d_pair_mc(ys + yq * (ybox >> 1) * 128, my, &full_bar[stage], ymask, kb * kBlockK,
                                yrow0 + yq * (ybox >> 1), kEvictLast);
          }
        } else if (ph1 ? a.ymn1 : a.ymn0) {
          for (int q = 0; q < (ybox >> 5); ++q) load(ys + q * 4096, my, yrow0 + q * 32, kb * kBlockK, kEvictLast);
        } else {
          load(ys, my, kb * kBlockK, yrow0, kEvictLast);
        }
        if (++stage == kNStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if constexpr (EPI == EPI_HUPDATE || EPI == EPI_RESID || EPI == EPI_KLQ || EPI == EPI_ABQ) {
      // (EPI_RESID / EPI_KLQ / EPI_ABQ: the staged tile is the 128-row x bn-column tile of V)
      if (a.h_prefetch && !helper) {
        // the ring is not used again: wait until every stage has been consumed, then reuse the
        // buffers for this CTA's 128-sample tile of the H master, [k][128 samples] fp32
        for (int s2 = 0; s2 < kNStages; ++s2) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == kNStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        mbar_arrive_expect_tx(&h_bar, static_cast<uint32_t>(bn) * 512u);
        for (int b = 0; b < (bn >> 5); ++b)
          tma_load_2d(sbase + b * 16384, &tmH, &h_bar, r0, n0 + b * 32, kEvictNormal);
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ------------------------------------------------ MMA issuer (leader CTA of a pair)
    int stage = 0;
    uint32_t phase = 0;
    for (int gch = 0; gch < n_items * nchunks; ++gch) {  // chunks of all items alternate between the two buffers
      const int ch = gch % nchunks;
      const int buf = gch & 1;
      const int use = gch >> 1;
      if (use > 0) {
        if constexpr (kPair) mbar_wait_cluster(&tempty_bar[buf], (use - 1) & 1);
        else mbar_wait(&tempty_bar[buf], (use - 1) & 1);
        tc_fence_after();
      }
      const bool ph1 = ch >= nchunk0;
      const int nkb = ph1 ? n1kb : min(chunk_kb, n0kb - ch * chunk_kb);
      const bool mn = ph1 ? (a.xmn1 != 0) : (a.xmn0 != 0);
      const bool ymn = ph1 ? (a.ymn1 != 0) : (a.ymn0 != 0);
      const uint32_t idesc = make_idesc_tf32(kTileM * TC::pair, bn, mn ? 1 : 0, ymn ? 1 : 0);
      const uint16_t pair_mask = static_cast<uint16_t>(0x3u << lead);           // this pair's two CTAs
      const uint16_t stage_mask = CG == 4 ? static_cast<uint16_t>(0xF) : pair_mask;  // everybody writing into the stage
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * kMaxN);
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t xs = sbase + stage * TC::stage_bytes;
        const uint32_t ys = xs + kStageBytesX;
#pragma unroll
        for (int s = 0; s < kBlockK / kUmmaK; ++s) {
          const uint64_t adesc = mn ? make_desc_mnmajor_sw128_32b(xs + s * 1024, 4096, 512)
                                    : make_desc_kmajor_sw128(xs + s * (kUmmaK * 4));
          const uint64_t bdesc = ymn ? make_desc_mnmajor_sw128_32b(ys + s * 1024, 4096, 512)
                                     : make_desc_kmajor_sw128(ys + s * (kUmmaK * 4));
          if constexpr (kPair) mma_tf32_ss_pair(d, adesc, bdesc, idesc, (i == 0 && s == 0) ? 0u : 1u);
          else mma_tf32_ss(d, adesc, bdesc, idesc, (i == 0 && s == 0) ? 0u : 1u);
        }
        // frees the smem slot (in both CTAs of a pair; CG = 4: one of the two arrivals in all four CTAs) when
        // these MMAs retire
        if constexpr (kPair) tc_commit_pair(&empty_bar[stage], stage_mask);
        else tc_commit(&empty_bar[stage]);
        if (++stage == kNStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      // chunk complete: wake the epilogue warps (of both CTAs)
      if constexpr (kPair) tc_commit_pair(&tfull_bar[buf], pair_mask);
      else tc_commit(&tfull_bar[buf]);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the 16-column groups
    const int row = r0 + q * 32 + lane;
    const bool row_ok = row < a.rows;
    const int ngroups = bn >> 4;
    const int g_begin = half ? ((ngroups + 1) >> 1) : 0;
    const int g_count = half ? (ngroups - g_begin) : ((ngroups + 1) >> 1);
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    float sum[kMaxGroups * 16];
#pragma unroll
    for (int i = 0; i < kMaxGroups * 16; ++i) sum[i] = 0.f;

    // EPI_KLQ / EPI_ABQ with a single accumulation chunk (K <= 512, the usual case): the epilogue
    // reads the accumulator straight from TMEM in a ROLLED loop over the column groups.  Unrolled
    // (as the register-resident sums require) it is ~8k straight-line instructions executed once per
    // CTA, and the kernel was instruction-fetch bound (stall_no_inst on 80 % of the samples).
    const bool direct = (EPI == EPI_KLQ || EPI == EPI_ABQ) && nchunk0 == 1 && n1kb == 0;
    if (direct) {
      mbar_wait(&tfull_bar[0], 0);
      tc_fence_after();
    }
    // promote finished phase-0 chunks from TMEM into registers
    int gch = 0;  // chunks seen so far (a helper pair runs through several row tiles)
    for (int item = 0; item < n_items; ++item) {
      for (int ch = 0; ch < (direct ? 0 : nchunk0); ++ch, ++gch) {
        const int buf = gch & 1;
        mbar_wait(&tfull_bar[buf], (gch >> 1) & 1);
        tc_fence_after();
        const uint32_t t0 = tlane + static_cast<uint32_t>(buf * kMaxN + g_begin * 16);
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) {
          if (g < g_count) {
            float v[16];
            tmem_ld16(t0 + g * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 16; ++t) sum[g * 16 + t] += v[t];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) mbar_arrive_remote(map_to_cta(smem_u32(&tempty_bar[buf]), lead));
          else mbar_arrive(&tempty_bar[buf]);
        }
      }
      if (helper) {
        // tail helper: the partial sum of this row tile goes to the scratch slab (coalesced along the rows, kept
        // in L2), then the flag its primary CTA is waiting for
        const int tile = first_tile + item * a.sk_helpers;
        float* p = a.sk_part + (static_cast<long long>(tile) * a.ncols + g_begin * 16) * (2 * kTileM) +
                   static_cast<int>(rank) * kTileM + q * 32 + lane;
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) {
          if (g < g_count) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              __stcg(p + (g * 16 + t) * (2 * kTileM), sum[g * 16 + t]);
              sum[g * 16 + t] = 0.f;
            }
          }
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // epilogue warps only
        if (warp == 2 && lane == 0) gate_publish(a.sk_flag + 2 * tile + static_cast<int>(rank), a.sk_epoch);
      }
    }
    if (!helper) {
    if (sk) {
      // primary: add the tail of the contraction, formed by a helper pair meanwhile
      if (lane == 0) gate_wait(a.sk_flag + 2 * pair_idx + static_cast<int>(rank), a.sk_epoch);
      __syncwarp();
      const float* p = a.sk_part + (static_cast<long long>(pair_idx) * a.ncols + g_begin * 16) * (2 * kTileM) +
                       static_cast<int>(rank) * kTileM + q * 32 + lane;
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        if (g < g_count) {
#pragma unroll
          for (int t = 0; t < 16; ++t) sum[g * 16 + t] += __ldcg(p + (g * 16 + t) * (2 * kTileM));
        }
      }
    }
    // second accumulator (if any) sits in the next buffer of the alternation
    const bool have1 = n1kb > 0;
    const uint32_t t1 =
        tlane + static_cast<uint32_t>((nchunk0 & 1) * kMaxN + g_begin * 16);
    if (have1) {
      mbar_wait(&tfull_bar[nchunk0 & 1], (nchunk0 >> 1) & 1);
      tc_fence_after();
    }
    const int col0 = n0 + g_begin * 16;

    if constexpr (EPI == EPI_STORE) {
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        if (g < g_count) {
          if (row_ok) {
            float* o = a.out0 + static_cast<long long>(split) * a.split_stride +
                       static_cast<long long>(col0 + g * 16) * a.ldo + row;
#pragma unroll
            for (int t = 0; t < 16; ++t) o[t * a.ldo] = sum[g * 16 + t];
          }
          if (a.nkb1 > 0 && split == 0) {
            float v[16];
            if (have1) {
              tmem_ld16(t1 + g * 16, v);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) v[t] = 0.f;
            }
            if (row_ok) {
              float* o = a.out1 + static_cast<long long>(col0 + g * 16) * a.ldo + row;
#pragma unroll
              for (int t = 0; t < 16; ++t) o[t * a.ldo] = v[t];
            }
          }
        }
      }
    } else {
      float s0 = 0.f, s1 = 0.f;  // per-thread partial sums (meaning depends on the epilogue)
      if constexpr (EPI == EPI_HUPDATE) {
        // fused multiplicative H update; thread = one sample (column of V / H), 16 basis rows at
        // a time.  Kept lean on purpose (fast division, warp-uniform branches hoisted): this code
        // runs after the last MMA with nothing left to hide behind.
        const bool vecD = a.dvec != nullptr;
        const bool staged = a.h_prefetch != 0;
        const bool write = !a.freeze && row_ok;
        const float* hsm = reinterpret_cast<const float*>(smem_raw + (sbase - smem_u32(smem_raw))) +
                           (g_begin * 16) * kTileM + q * 32 + lane;
        if (staged) mbar_wait(&h_bar, 0);
        // 32-bit element offsets from two 64-bit bases (one IMAD per address instead of a
        // 64-bit multiply-add chain); a tile spans at most 256 * ldh elements
        const long long hbase = static_cast<long long>(col0) * a.ldh + (row_ok ? row : 0);
        float* hm = a.Hm + hbase;
        float* hr = a.Hr32 + hbase;
        const unsigned ldh32 = static_cast<unsigned>(a.ldh);
        const float lam = a.lambda;
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) {
          if (g < g_count) {
            float dv[16];
            if (vecD) {
#pragma unroll
              for (int t = 0; t < 16; ++t) dv[t] = __ldg(a.dvec + col0 + g * 16 + t);
            } else if (have1) {
              tmem_ld16(t1 + g * 16, dv);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) dv[t] = 0.f;
            }
            float hv[16];
            if (staged) {
#pragma unroll
              for (int t = 0; t < 16; ++t) hv[t] = hsm[(g * 16 + t) * kTileM];
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) hv[t] = hm[(g * 16 + t) * ldh32];
            }
            if (!a.freeze) {
#pragma unroll
              for (int t = 0; t < 16; ++t)
                hv[t] = hv[t] * __fdividef(sum[g * 16 + t], fmaxf(dv[t] + lam, NMFB_EPS));
            }
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              dv[t] = tf32_rn(hv[t]);
              s0 = fmaf(sum[g * 16 + t], dv[t], s0);
              s1 += hv[t];
            }
            if (write) {
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                hm[(g * 16 + t) * ldh32] = hv[t];
                hr[(g * 16 + t) * ldh32] = dv[t];
              }
              if (a.Hc32 != nullptr) {
                float4* o = reinterpret_cast<float4*>(a.Hc32 + static_cast<long long>(row) * a.ldc +
                                                      (col0 + g * 16));
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  o[t] = make_float4(dv[4 * t], dv[4 * t + 1], dv[4 * t + 2], dv[4 * t + 3]);
              }
            }
          }
        }
      } else {
        // acc = tile of V_hat; thread = one row i of V, columns j of V in registers
        const int cols_ok = a.ncols_valid - col0;  // columns beyond the problem are padding
        const bool staged = (EPI == EPI_RESID || EPI == EPI_KLQ || EPI == EPI_ABQ) && a.h_prefetch != 0;
        const float* vsm = reinterpret_cast<const float*>(smem_raw + (sbase - smem_u32(smem_raw))) +
                           (g_begin * 16) * kTileM + q * 32 + lane;
        if (staged) mbar_wait(&h_bar, 0);  // V tile staged by TMA: [column][128 rows]
        if ((EPI == EPI_KLQ || EPI == EPI_ABQ) && direct) {
          const uint32_t ta = tlane + static_cast<uint32_t>(g_begin * 16);
          if constexpr (EPI == EPI_KLQ) {
            q_tile_direct<-1>(a, ta, g_count, col0, row, row_ok, cols_ok, staged, vsm, kTileM, s0, s1);
          } else {
            if (a.ab_mode == ABQ_IS)
              q_tile_direct<ABQ_IS>(a, ta, g_count, col0, row, row_ok, cols_ok, staged, vsm, kTileM, s0, s1);
            else if (a.ab_mode == ABQ_AB)
              q_tile_direct<ABQ_AB>(a, ta, g_count, col0, row, row_ok, cols_ok, staged, vsm, kTileM, s0, s1);
            else
              q_tile_direct<ABQ_AB_DUAL>(a, ta, g_count, col0, row, row_ok, cols_ok, staged, vsm, kTileM, s0, s1);
          }
        } else if constexpr (EPI == EPI_ABQ) {
          if (row_ok) {
            if (a.ab_mode == ABQ_IS)
              abq_tile<ABQ_IS, kMaxGroups>(a, sum, g_count, col0, row, cols_ok, staged, vsm, kTileM, s0);
            else if (a.ab_mode == ABQ_AB)
              abq_tile<ABQ_AB, kMaxGroups>(a, sum, g_count, col0, row, cols_ok, staged, vsm, kTileM, s0);
            else
              abq_tile<ABQ_AB_DUAL, kMaxGroups>(a, sum, g_count, col0, row, cols_ok, staged, vsm, kTileM, s0);
          }
        } else {
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) {
          if (g < g_count && row_ok) {
            const long long off = static_cast<long long>(col0 + g * 16) * a.ldv + row;
            if constexpr (EPI == EPI_RECON) {
#pragma unroll
              for (int t = 0; t < 16; ++t)
                if (g * 16 + t < cols_ok) a.Qout[off + t * a.ldv] = sum[g * 16 + t];
            } else {
              float v[16];
#pragma unroll
              for (int t = 0; t < 16; ++t)
                v[t] = (g * 16 + t < cols_ok)
                           ? (staged ? vsm[(g * 16 + t) * kTileM] : __ldg(a.Vsrc + off + t * a.ldv))
                           : 0.f;
              if constexpr (EPI == EPI_RESID) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                  const float d = v[t] - sum[g * 16 + t];
                  if (g * 16 + t < cols_ok) s0 = fmaf(d, d, s0);
                }
              } else if constexpr (EPI == EPI_KLQ) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                  if (g * 16 + t < cols_ok) {
                    const float sv = sum[g * 16 + t];
                    a.Qout[off + t * a.ldv] = tf32_rn(v[t] * fast_rcp(sv));
                    if (a.want_cost) {
                      s0 = fmaf(v[t], fast_lg2(sv), s0);  // log2: scaled by ln 2 below
                      s1 += sv;
                    }
                  }
                }
              }
            }
          }
        }
        }
        if constexpr (EPI == EPI_KLQ) s0 *= 0.6931471805599453f;
      }
      if (EPI != EPI_RECON && a.scal != nullptr) {
        double d0 = row_ok ? static_cast<double>(s0) : 0.0;
        double d1 = row_ok ? static_cast<double>(s1) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d0 += __shfl_xor_sync(0xffffffffu, d0, o);
          d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        }
        if (lane == 0) {
          red[warp - 2][0] = d0;
          red[warp - 2][1] = d1;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // epilogue warps only
        if (warp == 2 && lane == 0) {
          double p0 = 0.0, p1 = 0.0;
          for (int w = 0; w < kEpiWarps; ++w) {
            p0 += red[w][0];
            p1 += red[w][1];
          }
          atomicAdd(a.scal, p0);
          atomicAdd(a.scal + 1, p1);
        }
      }
    }
    }  // !helper
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();  // partner may still read our smem / TMEM
  if (warp == 1) {
    __syncwarp();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace nmfb
