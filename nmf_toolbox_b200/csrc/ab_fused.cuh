// ab_fused: the two-weight sibling of kl_fused for the Itakura-Saito and alpha-beta divergences
// (nmf.m:154-164 for the basis update, 185-195 for the encodings).  Per half iteration the reference forms
// V_hat = W*H, two element-wise weight matrices and two products with the other factor,
//
//     IS         Qn = V ./ V_hat.^2                 Qp = 1 ./ V_hat                 (nmf.m:155-156, 186-187)
//     AB         Qn = V.^a .* V_hat.^(b-1)          Qp = V_hat.^(a+b-1)             (nmf.m:162-163, 193-194)
//     AB, a = 0  Qn = V.^(a-1) .* V_hat.^b          Qp = V.^(a+b-1)                 (nmf.m:159-160, 190-191)
//     W half:  OUTn = Qn * H',  OUTp = Qp * H'      H half:  OUTn = (W' * Qn)',  OUTp = (W' * Qp)'
//
// Round 1 wrote Qn and Qp (two m x n matrices) to HBM and read each of them back twice.  Here, as in
// kl_fused.cuh, a tile of V_hat is formed in TMEM (MMA #1), sixteen epilogue warps turn it into the two
// weight tiles in place (TMEM -> registers -> TMEM, tf32 round-to-nearest; on request the sum of the
// per-element divergence, nmf.m:211-214), and two more MMAs with the A operand read from TMEM accumulate
// OUTn and OUTp - neither V_hat nor a weight matrix ever exists in HBM, V is streamed once per half.
//
// Layout, operands and rings are those of kl_fused (F = the factor whose rows are the output rows, G the
// other one, both [Kp][ld] with the long index contiguous; the H half reads a row-major copy of V).  TMEM:
// two S/Q slots of 128 columns ([0,64) S -> Qn, [64,128) Qp) and the two accumulators at [256, 256+Kp) and
// [384, 384+Kp).  The accumulators stay in TMEM for the whole work item: a promotion of both into registers
// (as kl_fused does against the tensor core's truncating accumulation) does not fit the register file, so
// the planner bounds a work item to kAbMaxTiles column tiles (512 accumulation steps: bias ~3e-5 relative,
// common to OUTn and OUTp, so it cancels in the update's ratio - measured at 8192^2, K = 128, IS: cost error
// against the float64 restatement 1.6e-7 / 1.7e-7 / 1.8e-7 for chains of 16 / 32 / 64 tiles); column splits do the rest.
#pragma once
#include "kl_fused.cuh"

namespace nmfb {

constexpr int kAbSlots = 2;        // S/Q slots in TMEM == ring of the MMA #1 view of G
constexpr int kAbMaxTiles = 64;    // column tiles per work item (accumulation chain of OUT: 8 steps per tile)
#ifndef NMFB_AB_VSLOTS
#define NMFB_AB_VSLOTS 2
#endif
constexpr int kAbVSlots = NMFB_AB_VSLOTS;  // ring of V tiles (a third slot fits, 225 KB in all, and changes nothing: 257 vs 258 us)
constexpr int kAbOffG2 = kKlOffG1 + kAbSlots * kKlG1Bytes;
constexpr int kAbOffV = kAbOffG2 + kKlG2Slots * kKlG2Bytes;
constexpr int kAbSmemBytes = kAbOffV + kAbVSlots * kKlVBytes + 1024;

struct AbArgs {
  int rows, cols, Kp;
  int tiles_per_split;
  int want_cost;
  float alpha, beta;
  float* out_n;    // partial slabs: out[split * slab + k * ldo + r]
  float* out_p;
  long long ldo;
  long long slab;
  double* scal;    // scal[0] += sum of the per-element divergence (want_cost)
  const int* stop;
};

// One element: the two weights and (optionally) its term of the divergence.  MODE as in panel_gemm.cuh.
template <int MODE>
__device__ __forceinline__ void ab_element(float vv, float sv, float al, float be, float rs, bool want_cost, float& qn,
                                           float& qp, float& s0) {
  if constexpr (MODE == ABQ_IS) {
    const float r = fast_rcp(sv);
    qp = r;
    qn = vv * r * r;
    if (want_cost) s0 += 0.6931471805599453f * fast_lg2(sv * fast_rcp(vv)) + vv * r - 1.f;  // nmf.m:212
  } else {  // powers as exp2(c log2 x); x.^0 == 1 (no 0 * inf)
    const float c0 = MODE == ABQ_AB ? al : al - 1.f, c1 = MODE == ABQ_AB ? be - 1.f : be;
    const float c2 = MODE == ABQ_AB ? 0.f : al + be - 1.f, c3 = MODE == ABQ_AB ? al + be - 1.f : 0.f;
    const float lv = fast_lg2(vv), ls = fast_lg2(sv);
    auto term = [](float c, float l) { return c == 0.f ? 0.f : c * l; };
    qn = fast_ex2(term(c0, lv) + term(c1, ls));
    qp = fast_ex2(term(c2, lv) + term(c3, ls));
    if (want_cost)  // nmf.m:214 (the prefactor is applied by the cost kernel)
      s0 += fast_ex2(term(al, lv) + term(be, ls)) -
            (al * fast_ex2(term(al + be, lv)) + be * fast_ex2(term(al + be, ls)) + be) * rs;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kKlThreads, 1)
ab_fused_kernel(const __grid_constant__ CUtensorMap tmF,    // F  [Kp][ld]  boxes 32 r x 32 k (MN-major, 32B-atom swizzle)
                const __grid_constant__ CUtensorMap tmG1,   // G  [Kp][ld]  boxes 32 c x 32 k (MN-major, 32B-atom swizzle)
                const __grid_constant__ CUtensorMap tmG2,   // G  [Kp][ld]  boxes 32 c x Kp/2 rows (K-major, 128B swizzle)
                const __grid_constant__ CUtensorMap tmV,    // VT [cols][ld] boxes 128 r x 64 c (linear)
                const AbArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t f_full;                 // leader: F tiles of both CTAs have landed
  __shared__ uint64_t g1_full[kAbSlots];      // leader: MMA #1 view of G (both CTAs' halves) landed
  __shared__ uint64_t g1_empty[kAbSlots];     // local : MMA #1 of the tile retired (commit, multicast)
  __shared__ uint64_t g2_full[kKlG2Slots];    // leader: MMA #2 view of G landed
  __shared__ uint64_t g2_empty[kKlG2Slots];   // local : both MMAs #2 of the tile retired (commit, multicast)
  __shared__ uint64_t v_full[kAbVSlots];      // local : this CTA's V tile landed
  __shared__ uint64_t v_empty[kAbVSlots];     // local : every epilogue warp has its V values in registers
  __shared__ uint64_t s_full[kAbSlots];       // local : S tile complete (commit, multicast)
  __shared__ uint64_t q_full[kAbSlots];       // leader: both CTAs' weight tiles are in TMEM
  __shared__ uint64_t sq_free[kAbSlots];      // leader: MMAs #2 of the tile retired, S/Q slot reusable
  __shared__ uint64_t o_full;                 // local : both accumulators complete (commit, multicast)
  __shared__ uint32_t tmem_slot;
  __shared__ double red[kKlEpiWarps];

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

  if (threadIdx.x == 0) {
    mbar_init(&f_full, 1);
    for (int i = 0; i < kAbSlots; ++i) {
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
    for (int i = 0; i < kAbVSlots; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], kKlEpiWarps);
    }
    mbar_init(&o_full, 1);
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
  const uint32_t tm_sq = tmem_base;          // slot s: [128 s, 128 s + 64) S -> Qn, [128 s + 64, 128 s + 128) Qp
  const uint32_t tm_out = tmem_base + 256;   // OUTn at +0, OUTp at +128

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
    int n1 = 0, n2 = 0, nv = 0;
    long long spin0 = clock64();
    while (n1 < ntiles || n2 < ntiles || nv < ntiles) {
      bool progress = false;
      if (n1 < ntiles && mbar_try_wait(&g1_empty[n1 % kAbSlots], ((n1 / kAbSlots) & 1) ^ 1)) {
        const int slot = n1 % kAbSlots;
        const int c0 = (t_begin + n1) * kKlTileC;
        if (rank == 0) mbar_arrive_expect_tx(&g1_full[slot], 2u * static_cast<uint32_t>(nkb) * 4096u);
        const uint32_t gb = map_to_cta(smem_u32(&g1_full[slot]), 0);
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d_pair(sbase + kKlOffG1 + slot * kKlG1Bytes + kb * 4096, &tmG1, gb,
                           c0 + static_cast<int>(rank) * 32, kb * 32, kEvictLast);
        ++n1;
        progress = true;
      }
      if (nv < ntiles && mbar_try_wait(&v_empty[nv % kAbVSlots], ((nv / kAbVSlots) & 1) ^ 1)) {
        const int slot = nv % kAbVSlots;
        mbar_arrive_expect_tx(&v_full[slot], kKlVBytes);
        tma_load_2d(sbase + kAbOffV + slot * kKlVBytes, &tmV, &v_full[slot], r0, (t_begin + nv) * kKlTileC,
                    kEvictFirst);
        ++nv;
        progress = true;
      }
      if (n2 < ntiles && mbar_try_wait(&g2_empty[n2 % kKlG2Slots], ((n2 / kKlG2Slots) & 1) ^ 1)) {
        const int slot = n2 % kKlG2Slots;
        const int c0 = (t_begin + n2) * kKlTileC;
        if (rank == 0) mbar_arrive_expect_tx(&g2_full[slot], 4u * static_cast<uint32_t>(Kp / 2) * 128u);
        const uint32_t gb = map_to_cta(smem_u32(&g2_full[slot]), 0);
        for (int cb = 0; cb < 2; ++cb)
          tma_load_2d_pair(sbase + kAbOffG2 + slot * kKlG2Bytes + cb * (Kp / 2) * 128, &tmG2, gb, c0 + cb * 32,
                           static_cast<int>(rank) * (Kp / 2), kEvictLast);
        ++n2;
        progress = true;
      }
      if (progress) {
        spin0 = clock64();
      } else if (clock64() - spin0 > 4000000000LL) {
        printf("nmfb: ab_fused producer timeout (block %d,%d)\n", blockIdx.x, blockIdx.y);
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
      const int slot = t % kAbSlots;
      const uint32_t use = static_cast<uint32_t>(t / kAbSlots);
      if (use > 0) mbar_wait(&sq_free[slot], (use - 1) & 1);  // MMAs #2 of tile t-2 have read this slot
      mbar_wait(&g1_full[slot], use & 1);
      tc_fence_after();
      const uint64_t bdesc0 = make_desc_mnmajor_sw128_32b(sbase + kKlOffG1 + slot * kKlG1Bytes, 4096, 512);
      const uint32_t d = tm_sq + static_cast<uint32_t>(slot * 128);
#pragma unroll
      for (int kb = 0; kb < kKlMaxKp / 32; ++kb) {
        if (kb < nkb) {
#pragma unroll
          for (int s2 = 0; s2 < kBlockK / kUmmaK; ++s2)
            mma_tf32_ss_pair(d, adesc0 + static_cast<uint64_t>((kb * 16384 + s2 * 1024) >> 4),
                             bdesc0 + static_cast<uint64_t>((kb * 4096 + s2 * 1024) >> 4), idesc1,
                             (kb == 0 && s2 == 0) ? 0u : 1u);
        }
      }
      tc_commit_pair(&s_full[slot], 0x3);
      tc_commit_pair(&g1_empty[slot], 0x3);
    }
  } else if (warp == 2 + kKlEpiWarps && lane == 0 && rank == 0) {
    // ------------------------------------------------ MMA #2 issuer: OUTn += Qn_t * G_t, OUTp += Qp_t * G_t
    const uint32_t idesc2 = make_idesc_tf32(2 * kTileM, Kp, 0, 0);  // A (= Q) from TMEM, G K-major
    const uint32_t cb_step = static_cast<uint32_t>((Kp / 2) * 128) >> 4;
    for (int t = 0; t < ntiles; ++t) {
      const int slot = t % kAbSlots;
      const int g2slot = t % kKlG2Slots;
      mbar_wait(&g2_full[g2slot], (t / kKlG2Slots) & 1);
      mbar_wait_cluster(&q_full[slot], (t / kAbSlots) & 1);
      tc_fence_after();
      const uint64_t bdesc0 = make_desc_kmajor_sw128(sbase + kAbOffG2 + g2slot * kKlG2Bytes);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const uint32_t d = tm_out + static_cast<uint32_t>(w * 128);
        const uint32_t qa = tm_sq + static_cast<uint32_t>(slot * 128 + w * kKlTileC);
#pragma unroll
        for (int s2 = 0; s2 < kKlTileC / kUmmaK; ++s2)  // 8 steps of 8 columns
          mma_tf32_ts_pair(d, qa + static_cast<uint32_t>(s2 * kUmmaK),
                           bdesc0 + static_cast<uint64_t>((s2 >> 2) * cb_step + (s2 & 3) * 2), idesc2,
                           (t == 0 && s2 == 0) ? 0u : 1u);
      }
      tc_commit_pair(&g2_empty[g2slot], 0x3);
      tc_commit_pair(&sq_free[slot], 0x1);
      if (t == ntiles - 1) tc_commit_pair(&o_full, 0x3);
    }
  } else if (warp >= 2 && warp < 2 + kKlEpiWarps) {
    // ------------------------------------------------ epilogue warps
    const int q = warp & 3;           // TMEM lane quarter
    const int sub = (warp - 2) >> 2;  // which 16 of the 64 tile columns / which quarter of the OUT columns
    const int row = r0 + q * 32 + lane;
    const bool row_ok = row < a.rows;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float al = a.alpha, be = a.beta;
    const float rs = 1.f / (al + be);  // alpha + beta == 0: Inf, as the reference's division
    const bool want_cost = a.want_cost != 0;
    float cs0 = 0.f;
    const float* smem_f = reinterpret_cast<const float*>(smem_raw + (sbase - smem_u32(smem_raw)));

    for (int t = 0; t < ntiles; ++t) {
      const int slot = t % kAbSlots;
      const uint32_t par = (t / kAbSlots) & 1;
      const int c0 = (t_begin + t) * kKlTileC + sub * 16;
      float va[16];
      {
        const int vslot = t % kAbVSlots;
        mbar_wait(&v_full[vslot], (t / kAbVSlots) & 1);
        const float* vt = smem_f + (kAbOffV + vslot * kKlVBytes) / 4 + (sub * 16) * kTileM + q * 32 + lane;
#pragma unroll
        for (int j = 0; j < 16; ++j) va[j] = vt[j * kTileM];
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_empty[vslot]);
      }
      mbar_wait(&s_full[slot], par);
      tc_fence_after();
      const uint32_t ta = tm_sq + lane_off + static_cast<uint32_t>(slot * 128 + sub * 16);
      float sv[16], qp[16];
      tmem_ld16(ta, sv);
      tmem_ld_wait();
      const bool full = row_ok && (c0 + 16 <= a.cols);
      if (full) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float qn;
          ab_element<MODE>(va[j], sv[j], al, be, rs, want_cost, qn, qp[j], cs0);
          sv[j] = tf32_rn(qn);
          qp[j] = tf32_rn(qp[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool ok = row_ok && (c0 + j < a.cols);
          float qn = 0.f, c = 0.f;
          qp[j] = 0.f;
          if (ok) ab_element<MODE>(va[j], sv[j], al, be, rs, want_cost, qn, qp[j], c);
          cs0 += c;
          sv[j] = ok ? tf32_rn(qn) : 0.f;  // padding must not inject NaN into OUT
          qp[j] = ok ? tf32_rn(qp[j]) : 0.f;
        }
      }
      tmem_st16(ta, sv);
      tmem_st16(ta + kKlTileC, qp);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&q_full[slot]), 0));
    }
    // both accumulators -> the partial slabs of this column split
    if (ntiles > 0) {
      mbar_wait(&o_full, 0);
      tc_fence_after();
    }
    const int ocols = Kp >> 2;  // OUT columns per thread and accumulator: 8, 16, 24 or 32
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      float* base = (w == 0 ? a.out_n : a.out_p) + static_cast<long long>(blockIdx.y) * a.slab +
                    static_cast<long long>(sub * ocols) * a.ldo + row;
      const uint32_t to = tm_out + lane_off + static_cast<uint32_t>(w * 128 + sub * ocols);
#pragma unroll
      for (int g = 0; g < kKlMaxKp / 4 / 8; ++g) {
        if (g * 8 < ocols) {
          float ov[8];
          if (ntiles > 0) {
            tmem_ld8(to + g * 8, ov);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ov[i] = 0.f;
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) base[static_cast<long long>(g * 8 + i) * a.ldo] = ov[i];
          }
        }
      }
    }
    if (want_cost && a.scal != nullptr) {
      double d0 = cs0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      if (lane == 0) red[warp - 2] = d0;
      asm volatile("bar.sync 1, %0;" ::"n"(kKlEpiWarps * 32) : "memory");
      if (warp == 2 && lane == 0) {
        double p0 = 0.0;
        for (int w = 0; w < kKlEpiWarps; ++w) p0 += red[w];
        atomicAdd(a.scal, p0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace nmfb
