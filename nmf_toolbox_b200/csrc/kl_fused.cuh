// kl_fused: one "flash-style" tensor-core kernel for both halves of the KL-divergence
// multiplicative update of nmf.m (lines 152-153 and 183-184).  For a block of rows r and a
// range of columns c of the data matrix it computes, tile by tile and entirely on chip,
//
//     S[r, c]   = sum_k F[r, k] G[c, k]          (tile of V_hat; tcgen05 MMA #1 -> TMEM)
//     Q[r, c]   = V[r, c] / S[r, c]              (epilogue warps: TMEM -> registers -> TMEM,
//                                                 tf32 round-to-nearest; optional cost sums)
//     OUT[r, k] += sum_c Q[r, c] G[c, k]         (tcgen05 MMA #2, A operand read from TMEM)
//
// so that neither V_hat nor Q = V ./ V_hat (both m x n) ever exists in HBM:
//   W half (nmf.m:152):  rows = rows of V,    F = W,  G = H',  OUT = (V ./ V_hat) * H'
//   H half (nmf.m:183):  rows = columns of V, F = H', G = W,   OUT = (W' * (V ./ V_hat))'
// Both factors live as [Kp][ld] arrays with the long index contiguous (W column-major, H
// row-major), which makes F and the MMA-#1 view of G MN-major operands and the MMA-#2 view of
// G a K-major operand - the same kernel serves both halves; the H half reads a row-major copy
// of V so that "rows" are contiguous in both cases.
//
// CTA pair (cta_group::2, 256 rows per pair), 64-column tiles.  Three operand rings in shared
// memory with their own depth - G for MMA #1 (4 slots, freed as soon as S is formed), G for
// MMA #2 and the V tile (2 slots each, freed late) - and four S/Q buffers in TMEM, so that
// MMA #1 runs up to four tiles ahead of MMA #2 and the TMA latency of a refill is covered:
//   warp 0    TMA producer: F once; per tile both views of G (this CTA's halves) and the V tile;
//             the three rings are served in whatever order their slots become free
//   warp 1    MMA #1 issuer (leader CTA); warp 18: MMA #2 issuer.  The MMAs of this kernel are
//             short (32-64 cycles), a single issuing thread cannot keep the tensor pipe busy;
//             the two streams only meet at the S/Q buffers (mbarrier sq_free)
//   warps 2-17 epilogue: S -> Q per tile (the MUFU-heavy part: one reciprocal, and for the cost
//             one logarithm, per element); promotion of finished OUT chunks into registers
// OUT is accumulated in chunks of kKlOutChunk tiles (64 MMA steps) that ping-pong between two
// TMEM buffers and are promoted with round-to-nearest adds (the tensor core truncates when it
// accumulates, see panel_gemm.cuh).  Columns are split over gridDim.y; each split writes a
// partial slab that a small kernel sums (and, for the H half, turns into the H update).
#pragma once
#include "panel_gemm.cuh"

namespace nmfb {

constexpr int kKlTileC = 64;       // columns per tile
constexpr int kKlG1Slots = 4;      // ring of the MMA #1 view of G == number of S/Q buffers in TMEM
constexpr int kKlG2Slots = 2;      // ring of the MMA #2 view of G
constexpr int kKlVSlots = 2;       // ring of V tiles
#ifndef NMFB_KL_OUT_CHUNK
#define NMFB_KL_OUT_CHUNK 8
#endif
// tiles per OUT accumulation chunk: 8 * 8 = 64 MMA steps, the chain length of the panel GEMM's chunks (bias ~3e-6);
// 4 -> 8 is worth 2 % of config 3 (823 -> 840 it/s), the cost curve moves by 6e-9 relative
constexpr int kKlOutChunk = NMFB_KL_OUT_CHUNK;
constexpr int kKlMaxKp = 128;
constexpr int kKlEpiWarps = 16;    // four warps per TMEM lane quarter, 16 tile columns each
constexpr int kKlThreads = 64 + kKlEpiWarps * 32 + 32;  // + one more MMA-issuing warp
constexpr int kKlFBytes = kTileM * kKlMaxKp * 4;                 // F: 128 rows x Kp            64 KB
constexpr int kKlG1Bytes = (kKlTileC / 2) * kKlMaxKp * 4;        // 32 c x Kp (this CTA's half) 16 KB
constexpr int kKlG2Bytes = (kKlMaxKp / 2) * kKlTileC * 4;        // Kp/2 x 64 c                 16 KB
constexpr int kKlVBytes = kKlTileC * kTileM * 4;                 // 64 c x 128 r                32 KB
constexpr int kKlOffG1 = kKlFBytes;
constexpr int kKlOffG2 = kKlOffG1 + kKlG1Slots * kKlG1Bytes;
constexpr int kKlOffV = kKlOffG2 + kKlG2Slots * kKlG2Bytes;
constexpr int kKlSmemBytes = kKlOffV + kKlVSlots * kKlVBytes + 1024;  // 225 KB

struct KlArgs {
  int rows;        // rows of this problem (m for the W half, n for the H half)
  int cols;        // columns to contract over (n resp. m)
  int Kp;          // padded number of basis vectors: 32, 64, 96 or 128
  int tiles_per_split;
  int want_cost;
  float* out;      // partial slabs: out[split * slab + k * ldo + r]
  long long ldo;
  long long slab;
  double* scal;    // scal[0] += sum V .* log(V_hat), scal[1] += sum V_hat   (want_cost)
  const int* stop;
};

__device__ __forceinline__ void mma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kKlThreads, 1)
kl_fused_kernel(const __grid_constant__ CUtensorMap tmF,    // F  [Kp][ld]  boxes 32 r x 32 k (MN-major, 32B-atom swizzle)
                const __grid_constant__ CUtensorMap tmG1,   // G  [Kp][ld]  boxes 32 c x 32 k (MN-major, 32B-atom swizzle)
                const __grid_constant__ CUtensorMap tmG2,   // G  [Kp][ld]  boxes 32 c x Kp/2 rows (K-major, 128B swizzle)
                const __grid_constant__ CUtensorMap tmV,    // VT [cols][ld] boxes 128 r x 64 c (linear)
                const KlArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t f_full;                   // leader: F tiles of both CTAs have landed
  __shared__ uint64_t g1_full[kKlG1Slots];      // leader: MMA #1 view of G (both CTAs' halves) landed
  __shared__ uint64_t g1_empty[kKlG1Slots];     // local : MMA #1 of the tile retired (commit, multicast)
  __shared__ uint64_t g2_full[kKlG2Slots];      // leader: MMA #2 view of G landed
  __shared__ uint64_t g2_empty[kKlG2Slots];     // local : MMA #2 of the tile retired (commit, multicast)
  __shared__ uint64_t v_full[kKlVSlots];        // local : this CTA's V tile landed
  __shared__ uint64_t v_empty[kKlVSlots];       // local : every epilogue warp has its V values in registers
  __shared__ uint64_t s_full[kKlG1Slots];       // local : S tile complete (commit, multicast)
  __shared__ uint64_t q_full[kKlG1Slots];       // leader: both CTAs' Q tiles are in TMEM
  __shared__ uint64_t sq_free[kKlG1Slots];      // leader: MMA #2 of the tile retired, S/Q buffer reusable
  __shared__ uint64_t o_full[2];                // local : OUT chunk complete (commit, multicast)
  __shared__ uint64_t o_empty[2];               // leader: OUT buffer promoted by both CTAs
  __shared__ uint32_t tmem_slot;
  __shared__ double red[kKlEpiWarps][2];

  if (a.stop != nullptr && *a.stop != 0) return;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t f_smem = sbase;
  const int Kp = a.Kp;
  const int nkb = Kp >> 5;  // k-blocks of 32
  const int r0 = static_cast<int>(blockIdx.x >> 1) * (2 * kTileM) + static_cast<int>(rank) * kTileM;
  const int total_tiles = (a.cols + kKlTileC - 1) / kKlTileC;
  const int t_begin = blockIdx.y * a.tiles_per_split;
  const int ntiles = max(0, min(total_tiles, t_begin + a.tiles_per_split) - t_begin);
  const int nchunks = (ntiles + kKlOutChunk - 1) / kKlOutChunk;

  if (threadIdx.x == 0) {
    mbar_init(&f_full, 1);
    for (int i = 0; i < kKlG1Slots; ++i) {
      mbar_init(&g1_full[i], 1);
      mbar_init(&g1_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&q_full[i], 2 * kKlEpiWarps);
      mbar_init(&sq_free[i], 1);
    }
    for (int i = 0; i < kKlG2Slots; ++i) {
      mbar_init(&g2_full[i], 1);
      mbar_init(&g2_empty[i], 1);
    }
    for (int i = 0; i < kKlVSlots; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], kKlEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&o_full[b], 1);
      mbar_init(&o_empty[b], 2 * kKlEpiWarps);
    }
    fence_barrier_init();
    prefetch_tmap(&tmF);
    prefetch_tmap(&tmG1);
    prefetch_tmap(&tmG2);
    prefetch_tmap(&tmV);
  }
  if (warp == 1) {
    tmem_alloc_pair(&tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // TMEM columns: four 64-column S/Q buffers at [0,256); OUT buffers at [256, 256+Kp) and [384, 384+Kp)
  const uint32_t tm_sq = tmem_base;
  const uint32_t tm_out = tmem_base + 256;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ TMA producer
    {  // F: this CTA's 128 rows, all k-blocks, loaded once; counted on the leader's barrier
      if (rank == 0) mbar_arrive_expect_tx(&f_full, 2u * static_cast<uint32_t>(nkb) * 16384u);
      const uint32_t fb = map_to_cta(smem_u32(&f_full), 0);
      for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          tma_load_2d_pair(f_smem + kb * 16384 + q * 4096, &tmF, fb, r0 + q * 32, kb * 32, kEvictLast);
    }
    // three independent rings, each refilled as soon as its next slot is free
    int n1 = 0, n2 = 0, nv = 0;
    long long spin0 = clock64();
    while (n1 < ntiles || n2 < ntiles || nv < ntiles) {
      bool progress = false;
      if (n1 < ntiles && mbar_try_wait(&g1_empty[n1 % kKlG1Slots], ((n1 / kKlG1Slots) & 1) ^ 1)) {
        const int slot = n1 % kKlG1Slots;
        const int c0 = (t_begin + n1) * kKlTileC;
        if (rank == 0) mbar_arrive_expect_tx(&g1_full[slot], 2u * static_cast<uint32_t>(nkb) * 4096u);
        const uint32_t gb = map_to_cta(smem_u32(&g1_full[slot]), 0);
        // this CTA's 32 columns of the tile, one 32 c x 32 k box per k-block
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d_pair(sbase + kKlOffG1 + slot * kKlG1Bytes + kb * 4096, &tmG1, gb,
                           c0 + static_cast<int>(rank) * 32, kb * 32, kEvictLast);
        ++n1;
        progress = true;
      }
      if (nv < ntiles && mbar_try_wait(&v_empty[nv % kKlVSlots], ((nv / kKlVSlots) & 1) ^ 1)) {
        const int slot = nv % kKlVSlots;
        mbar_arrive_expect_tx(&v_full[slot], kKlVBytes);
        tma_load_2d(sbase + kKlOffV + slot * kKlVBytes, &tmV, &v_full[slot], r0, (t_begin + nv) * kKlTileC,
                    kEvictFirst);
        ++nv;
        progress = true;
      }
      if (n2 < ntiles && mbar_try_wait(&g2_empty[n2 % kKlG2Slots], ((n2 / kKlG2Slots) & 1) ^ 1)) {
        const int slot = n2 % kKlG2Slots;
        const int c0 = (t_begin + n2) * kKlTileC;
        if (rank == 0) mbar_arrive_expect_tx(&g2_full[slot], 4u * static_cast<uint32_t>(Kp / 2) * 128u);
        const uint32_t gb = map_to_cta(smem_u32(&g2_full[slot]), 0);
        // this CTA's Kp/2 basis rows, two boxes of 32 c
        for (int cb = 0; cb < 2; ++cb)
          tma_load_2d_pair(sbase + kKlOffG2 + slot * kKlG2Bytes + cb * (Kp / 2) * 128, &tmG2, gb, c0 + cb * 32,
                           static_cast<int>(rank) * (Kp / 2), kEvictLast);
        ++n2;
        progress = true;
      }
      if (progress) {
        spin0 = clock64();
      } else if (clock64() - spin0 > 4000000000LL) {
        printf("nmfb: kl_fused producer timeout (block %d,%d)\n", blockIdx.x, blockIdx.y);
        __trap();
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ------------------------------------------------ MMA #1 issuer: S_t = F * G_t'
    const uint32_t idesc1 = make_idesc_tf32(2 * kTileM, kKlTileC, 1, 1);  // both operands MN-major
    const uint64_t adesc0 = make_desc_mnmajor_sw128_32b(f_smem, 4096, 512);
    mbar_wait(&f_full, 0);
    tc_fence_after();
    for (int t = 0; t < ntiles; ++t) {
      const int slot = t % kKlG1Slots;
      const uint32_t use = static_cast<uint32_t>(t / kKlG1Slots);
      if (use > 0) mbar_wait(&sq_free[slot], (use - 1) & 1);  // MMA #2 of tile t-4 has read this buffer
      mbar_wait(&g1_full[slot], use & 1);
      tc_fence_after();
      const uint64_t bdesc0 = make_desc_mnmajor_sw128_32b(sbase + kKlOffG1 + slot * kKlG1Bytes, 4096, 512);
      const uint32_t d = tm_sq + static_cast<uint32_t>(slot * kKlTileC);
#pragma unroll
      for (int kb = 0; kb < kKlMaxKp / 32; ++kb) {
        if (kb < nkb) {
#pragma unroll
          for (int s2 = 0; s2 < kBlockK / kUmmaK; ++s2)  // descriptors differ only in the 16-byte address field
            mma_tf32_ss_pair(d, adesc0 + static_cast<uint64_t>((kb * 16384 + s2 * 1024) >> 4),
                             bdesc0 + static_cast<uint64_t>((kb * 4096 + s2 * 1024) >> 4), idesc1,
                             (kb == 0 && s2 == 0) ? 0u : 1u);
        }
      }
      tc_commit_pair(&s_full[slot], 0x3);
      tc_commit_pair(&g1_empty[slot], 0x3);
    }
  } else if (warp == 2 + kKlEpiWarps && lane == 0 && rank == 0) {
    // ------------------------------------------------ MMA #2 issuer: OUT[chunk buffer] += Q_t * G_t
    const uint32_t idesc2 = make_idesc_tf32(2 * kTileM, Kp, 0, 0);  // A (= Q) from TMEM, G K-major
    const uint32_t cb_step = static_cast<uint32_t>((Kp / 2) * 128) >> 4;
    for (int t = 0; t < ntiles; ++t) {
      const int slot = t % kKlG1Slots;
      const int g2slot = t % kKlG2Slots;
      const int ch = t / kKlOutChunk;
      const int buf = ch & 1;
      const bool first = (t % kKlOutChunk) == 0;
      if (first && ch >= 2) mbar_wait_cluster(&o_empty[buf], ((ch >> 1) - 1) & 1);  // chunk ch-2 promoted
      mbar_wait(&g2_full[g2slot], (t / kKlG2Slots) & 1);
      mbar_wait_cluster(&q_full[slot], (t / kKlG1Slots) & 1);
      tc_fence_after();
      const uint64_t bdesc0 = make_desc_kmajor_sw128(sbase + kKlOffG2 + g2slot * kKlG2Bytes);
      const uint32_t d = tm_out + static_cast<uint32_t>(buf * 128);
      const uint32_t qa = tm_sq + static_cast<uint32_t>(slot * kKlTileC);
#pragma unroll
      for (int s2 = 0; s2 < kKlTileC / kUmmaK; ++s2)  // 8 steps of 8 columns
        mma_tf32_ts_pair(d, qa + static_cast<uint32_t>(s2 * kUmmaK),
                         bdesc0 + static_cast<uint64_t>((s2 >> 2) * cb_step + (s2 & 3) * 2), idesc2,
                         (first && s2 == 0) ? 0u : 1u);
      tc_commit_pair(&g2_empty[g2slot], 0x3);
      tc_commit_pair(&sq_free[slot], 0x1);
      const bool last_of_chunk = ((t % kKlOutChunk) == kKlOutChunk - 1) || (t == ntiles - 1);
      if (last_of_chunk) tc_commit_pair(&o_full[buf], 0x3);
    }
  } else if (warp >= 2 && warp < 2 + kKlEpiWarps) {
    // ------------------------------------------------ epilogue warps
    const int q = warp & 3;           // TMEM lane quarter
    const int sub = (warp - 2) >> 2;  // which 16 of the 64 tile columns / which quarter of the OUT columns
    const int row = r0 + q * 32 + lane;
    const bool row_ok = row < a.rows;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int ocols = Kp >> 2;  // OUT columns per thread: 8, 16, 24 or 32
    float osum[kKlMaxKp / 4];
#pragma unroll
    for (int i = 0; i < kKlMaxKp / 4; ++i) osum[i] = 0.f;
    float cs0 = 0.f, cs1 = 0.f;
    const float* smem_f = reinterpret_cast<const float*>(smem_raw + (sbase - smem_u32(smem_raw)));
    auto promote = [&](int buf) {
      const uint32_t to = tm_out + lane_off + static_cast<uint32_t>(buf * 128 + sub * ocols);
#pragma unroll
      for (int g = 0; g < kKlMaxKp / 4 / 8; ++g) {
        if (g * 8 < ocols) {
          float ov[8];
          tmem_ld8(to + g * 8, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) osum[g * 8 + i] += ov[i];
        }
      }
    };

    for (int t = 0; t < ntiles; ++t) {
      const int stage = t % kKlG1Slots;
      const uint32_t par = (t / kKlG1Slots) & 1;
      const int c0 = (t_begin + t) * kKlTileC + sub * 16;
      // this thread's 16 values of V: row `row`, columns c0 .. c0+15, from the staged tile [c][128 r]
      float va[16];
      {
        const int vslot = t % kKlVSlots;
        mbar_wait(&v_full[vslot], (t / kKlVSlots) & 1);
        const float* vt = smem_f + (kKlOffV + vslot * kKlVBytes) / 4 + (sub * 16) * kTileM + q * 32 + lane;
#pragma unroll
        for (int j = 0; j < 16; ++j) va[j] = vt[j * kTileM];
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_empty[vslot]);  // slot can be refilled while we work on the tile
      }
      mbar_wait(&s_full[stage], par);
      tc_fence_after();
      const uint32_t ta = tm_sq + lane_off + static_cast<uint32_t>(stage * kKlTileC + sub * 16);
      float sv[16];
      tmem_ld16(ta, sv);
      tmem_ld_wait();
      // fast path: whole tile in range (every tile but possibly the last one of the matrix)
      const bool full = row_ok && (c0 + 16 <= a.cols);
      if (full) {
        if (a.want_cost) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s = sv[j];
            cs0 = fmaf(va[j], fast_lg2(s), cs0);  // log2 here, scaled by ln 2 once at the end
            cs1 += s;
            sv[j] = tf32_rn(va[j] * fast_rcp(s));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) sv[j] = tf32_rn(va[j] * fast_rcp(sv[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float s = sv[j];
          const bool ok = row_ok && (c0 + j < a.cols);
          if (a.want_cost && ok) {
            cs0 = fmaf(va[j], fast_lg2(s), cs0);
            cs1 += s;
          }
          sv[j] = ok ? tf32_rn(va[j] * fast_rcp(s)) : 0.f;  // padding must not inject NaN into OUT
        }
      }
      tmem_st16(ta, sv);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&q_full[stage]), 0));
      // promote the OUT chunk closed by tile t-1 (its MMA #2 is issued right after MMA #1 of
      // tile t, which has completed since s_full[t] fired)
      if (t >= 1 && ((t - 1) % kKlOutChunk) == kKlOutChunk - 1) {
        const int ch = (t - 1) / kKlOutChunk;
        const int buf = ch & 1;
        mbar_wait(&o_full[buf], (ch >> 1) & 1);
        tc_fence_after();
        promote(buf);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&o_empty[buf]), 0));
      }
    }
    if (ntiles > 0) {  // the chunk that ends with the last tile (complete or partial)
      const int ch = (ntiles - 1) / kKlOutChunk;
      const int buf = ch & 1;
      mbar_wait(&o_full[buf], (ch >> 1) & 1);
      tc_fence_after();
      promote(buf);
    }
    // partial slab of this column split
    if (row_ok) {
      float* o = a.out + static_cast<long long>(blockIdx.y) * a.slab + static_cast<long long>(sub * ocols) * a.ldo + row;
#pragma unroll
      for (int i = 0; i < kKlMaxKp / 4; ++i)
        if (i < ocols) o[static_cast<long long>(i) * a.ldo] = osum[i];
    }
    if (a.want_cost && a.scal != nullptr) {
      double d0 = static_cast<double>(cs0) * 0.6931471805599453, d1 = cs1;  // log2 -> ln
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      }
      if (lane == 0) {
        red[warp - 2][0] = d0;
        red[warp - 2][1] = d1;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kKlEpiWarps * 32) : "memory");
      if (warp == 2 && lane == 0) {
        double p0 = 0.0, p1 = 0.0;
        for (int w = 0; w < kKlEpiWarps; ++w) {
          p0 += red[w][0];
          p1 += red[w][1];
        }
        atomicAdd(a.scal, p0);
        atomicAdd(a.scal + 1, p1);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
  (void)nchunks;
}

}  // namespace nmfb
