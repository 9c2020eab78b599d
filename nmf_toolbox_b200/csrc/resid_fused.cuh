// resid_fused: the objective 0.5*|V - W*H|_F^2 of nmfsc.m (lines 139, 161, 212, 238) in one
// streaming tensor-core kernel.  The line search of nmfsc compares objectives that differ by
// parts in 1e5, so V_hat is formed at fp32 accuracy as a split-tf32 product
//     S = W_hi H_hi + W_lo H_hi + W_hi H_lo        (three MMA segments per tile)
// and never leaves the chip: a CTA pair (cta_group::2) keeps its 256 rows of W (head and tail)
// resident in shared memory, streams 64-column tiles of H (head and tail, 2-slot ring) and of V
// (TMA, 1 slot), accumulates S in one of four TMEM buffers and 16 epilogue warps reduce
// (V - S)^2 straight from TMEM.  Same skeleton as kl_fused.cuh without the second MMA.
#pragma once
#include "ew_kernels.cuh"
#include "kl_fused.cuh"

namespace nmfb {

constexpr int kRsSBufs = 4;                                         // S buffers in TMEM
constexpr int kRsGSlots = 2;                                        // ring of H tiles (head + tail per slot)
constexpr int kRsGBytes = 2 * kKlG1Bytes;                           // 32 KB
constexpr int kRsOffFlo = kKlFBytes;                                // W tail behind the W head
constexpr int kRsOffG = 2 * kKlFBytes;
constexpr int kRsOffV = kRsOffG + kRsGSlots * kRsGBytes;
constexpr int kRsSmemBytes = kRsOffV + kKlVBytes + 1024;            // 64 + 64 + 64 + 32 KB

struct ResidArgs {
  int rows, cols, Kp;
  int tiles_per_split;
  double* scal;  // scal[0] += sum (V - S)^2
  const int* skip;  // device flag: non-zero = do nothing (phase guard of the device-side line search)
  LsFin fin;        // what the last block does with the finished sum (ew_kernels.cuh: ls_finish)
};

__global__ void __launch_bounds__(64 + kKlEpiWarps * 32, 1)
resid_fused_kernel(const __grid_constant__ CUtensorMap tmFhi, const __grid_constant__ CUtensorMap tmFlo,
                   const __grid_constant__ CUtensorMap tmGhi, const __grid_constant__ CUtensorMap tmGlo,
                   const __grid_constant__ CUtensorMap tmV, const ResidArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t f_full;               // leader: W head + tail of both CTAs landed
  __shared__ uint64_t g_full[kRsGSlots];    // leader: H tile (head + tail, both CTAs' halves) landed
  __shared__ uint64_t g_empty[kRsGSlots];   // local : the MMAs of the tile retired (commit, multicast)
  __shared__ uint64_t v_full, v_empty;      // local : V tile landed / copied to registers by all warps
  __shared__ uint64_t s_full[kRsSBufs];     // local : S tile complete (commit, multicast)
  __shared__ uint64_t s_free[kRsSBufs];     // leader: both CTAs' epilogues have read the S buffer
  __shared__ uint32_t tmem_slot;
  __shared__ double red[kKlEpiWarps];

  if (a.skip != nullptr && *a.skip != 0) return;  // uniform across the grid
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int Kp = a.Kp;
  const int nkb = Kp >> 5;
  const int r0 = static_cast<int>(blockIdx.x >> 1) * (2 * kTileM) + static_cast<int>(rank) * kTileM;
  const int total_tiles = (a.cols + kKlTileC - 1) / kKlTileC;
  const int t_begin = blockIdx.y * a.tiles_per_split;
  const int ntiles = max(0, min(total_tiles, t_begin + a.tiles_per_split) - t_begin);

  if (threadIdx.x == 0) {
    mbar_init(&f_full, 1);
    for (int i = 0; i < kRsGSlots; ++i) {
      mbar_init(&g_full[i], 1);
      mbar_init(&g_empty[i], 1);
    }
    mbar_init(&v_full, 1);
    mbar_init(&v_empty, kKlEpiWarps);
    for (int i = 0; i < kRsSBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 2 * kKlEpiWarps);
    }
    fence_barrier_init();
    prefetch_tmap(&tmFhi);
    prefetch_tmap(&tmFlo);
    prefetch_tmap(&tmGhi);
    prefetch_tmap(&tmGlo);
    prefetch_tmap(&tmV);
  }
  if (warp == 1) {
    tmem_alloc_pair(&tmem_slot, 256);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------ TMA producer
    {
      if (rank == 0) mbar_arrive_expect_tx(&f_full, 4u * static_cast<uint32_t>(nkb) * 16384u);
      const uint32_t fb = map_to_cta(smem_u32(&f_full), 0);
      for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tma_load_2d_pair(sbase + kb * 16384 + q * 4096, &tmFhi, fb, r0 + q * 32, kb * 32, kEvictLast);
          tma_load_2d_pair(sbase + kRsOffFlo + kb * 16384 + q * 4096, &tmFlo, fb, r0 + q * 32, kb * 32, kEvictLast);
        }
    }
    int ng = 0, nv = 0;
    long long spin0 = clock64();
    while (ng < ntiles || nv < ntiles) {
      bool progress = false;
      if (ng < ntiles && mbar_try_wait(&g_empty[ng % kRsGSlots], ((ng / kRsGSlots) & 1) ^ 1)) {
        const int slot = ng % kRsGSlots;
        const int c0 = (t_begin + ng) * kKlTileC + static_cast<int>(rank) * 32;
        if (rank == 0) mbar_arrive_expect_tx(&g_full[slot], 4u * static_cast<uint32_t>(nkb) * 4096u);
        const uint32_t gb = map_to_cta(smem_u32(&g_full[slot]), 0);
        const uint32_t base = sbase + kRsOffG + slot * kRsGBytes;
        for (int kb = 0; kb < nkb; ++kb) {
          tma_load_2d_pair(base + kb * 4096, &tmGhi, gb, c0, kb * 32, kEvictLast);
          tma_load_2d_pair(base + kKlG1Bytes + kb * 4096, &tmGlo, gb, c0, kb * 32, kEvictLast);
        }
        ++ng;
        progress = true;
      }
      if (nv < ntiles && mbar_try_wait(&v_empty, (nv & 1) ^ 1)) {
        mbar_arrive_expect_tx(&v_full, kKlVBytes);
        tma_load_2d(sbase + kRsOffV, &tmV, &v_full, r0, (t_begin + nv) * kKlTileC, kEvictNormal);
        ++nv;
        progress = true;
      }
      if (progress) {
        spin0 = clock64();
      } else if (clock64() - spin0 > 4000000000LL) {
        printf("nmfb: resid_fused producer timeout (block %d,%d)\n", blockIdx.x, blockIdx.y);
        __trap();
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ------------------------------------------------ MMA issuer: S_t = W_hi H_hi' + W_lo H_hi' + W_hi H_lo'
    const uint32_t idesc = make_idesc_tf32(2 * kTileM, kKlTileC, 1, 1);
    const uint64_t ahi = make_desc_mnmajor_sw128_32b(sbase, 4096, 512);
    const uint64_t alo = make_desc_mnmajor_sw128_32b(sbase + kRsOffFlo, 4096, 512);
    mbar_wait(&f_full, 0);
    tc_fence_after();
    for (int t = 0; t < ntiles; ++t) {
      const int sb = t % kRsSBufs;
      const int slot = t % kRsGSlots;
      const uint32_t use = static_cast<uint32_t>(t / kRsSBufs);
      if (use > 0) mbar_wait_cluster(&s_free[sb], (use - 1) & 1);
      mbar_wait(&g_full[slot], (t / kRsGSlots) & 1);
      tc_fence_after();
      const uint64_t bhi = make_desc_mnmajor_sw128_32b(sbase + kRsOffG + slot * kRsGBytes, 4096, 512);
      const uint64_t blo = make_desc_mnmajor_sw128_32b(sbase + kRsOffG + slot * kRsGBytes + kKlG1Bytes, 4096, 512);
      const uint32_t d = tmem_base + static_cast<uint32_t>(sb * kKlTileC);
#pragma unroll
      for (int seg = 0; seg < 3; ++seg) {
        const uint64_t a0 = seg == 1 ? alo : ahi;
        const uint64_t b0 = seg == 2 ? blo : bhi;
#pragma unroll
        for (int kb = 0; kb < kKlMaxKp / 32; ++kb) {
          if (kb < nkb) {
#pragma unroll
            for (int s2 = 0; s2 < kBlockK / kUmmaK; ++s2)
              mma_tf32_ss_pair(d, a0 + static_cast<uint64_t>((kb * 16384 + s2 * 1024) >> 4),
                               b0 + static_cast<uint64_t>((kb * 4096 + s2 * 1024) >> 4), idesc,
                               (seg == 0 && kb == 0 && s2 == 0) ? 0u : 1u);
          }
        }
      }
      tc_commit_pair(&s_full[sb], 0x3);
      tc_commit_pair(&g_empty[slot], 0x3);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;  // 16 of the 64 tile columns
    const int row = r0 + q * 32 + lane;
    const bool row_ok = row < a.rows;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float* vt = reinterpret_cast<const float*>(smem_raw + (sbase - smem_u32(smem_raw))) + kRsOffV / 4 +
                      (sub * 16) * kTileM + q * 32 + lane;
    float acc = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int sb = t % kRsSBufs;
      const int c0 = (t_begin + t) * kKlTileC + sub * 16;
      float va[16];
      mbar_wait(&v_full, t & 1);
#pragma unroll
      for (int j = 0; j < 16; ++j) va[j] = vt[j * kTileM];
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_empty);
      mbar_wait(&s_full[sb], (t / kRsSBufs) & 1);
      tc_fence_after();
      float sv[16];
      tmem_ld16(tmem_base + lane_off + static_cast<uint32_t>(sb * kKlTileC + sub * 16), sv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&s_free[sb]), 0));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float dlt = va[j] - sv[j];
        if (row_ok && (c0 + j < a.cols)) acc = fmaf(dlt, dlt, acc);
      }
    }
    double d0 = acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    if (lane == 0) red[warp - 2] = d0;
    asm volatile("bar.sync 1, %0;" ::"n"(kKlEpiWarps * 32) : "memory");
    if (warp == 2 && lane == 0) {
      double p = 0.0;
      for (int w = 0; w < kKlEpiWarps; ++w) p += red[w];
      atomicAdd(a.scal, p);
      if (a.fin.mode != LSFIN_NONE) {  // last block: the line-search decision / cost entry on the complete sum
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(a.fin.ticket, 1u) == total - 1) {
          *a.fin.ticket = 0u;
          __threadfence();
          ls_finish(a.fin, a.scal);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair(tmem_base, 256);
  }
}

}  // namespace nmfb
