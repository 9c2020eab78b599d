// Host-side helpers for panel_gemm: TMA tensor-map construction (through the
// driver entry point, so the library does not link libcuda) and launching.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "panel_gemm.cuh"

namespace nmfb {

// A 2-D fp32 matrix as TMA sees it: `inner` contiguous elements per row,
// `outer` rows, `pitch` elements between rows (pitch*4 must be a multiple of 16).
struct Mat2D {
  const float* base;
  long long inner;
  long long outer;
  long long pitch;
};

// Operand description for one phase of panel_gemm.
struct GemmOperand {
  Mat2D m;
  bool mn_major;  // the operand's rows (output rows for X, output columns for Y) are contiguous
};

// Returns an empty string on success, else an error message.
// atom32b selects the 128B-swizzle-with-32B-atom mode that MN-major fp32 operands need.
std::string make_tmap(CUtensorMap* out, const Mat2D& m, uint32_t box_inner, uint32_t box_outer,
                      bool atom32b, bool no_swizzle = false);

// Number of phase-0 k-blocks each split handles so that tiles*chunks*splits ~ fills the GPU.
int choose_splits(int tiles, int nkb0, int num_sms, int* kb_per_split);

struct GemmLaunch {
  CUtensorMap tmX0, tmY0, tmX1, tmY1;
  CUtensorMap tmXb, tmYb, tmXc, tmYc;  // extra phase-0 segments (same shapes / majorness as X0, Y0)
  CUtensorMap tmH;                     // EPI_HUPDATE: H master tile (set_h_prefetch)
  GemmArgs args;
  dim3 grid;
  int cg = 1;  // 2 = CTA-pair kernel (256-row tiles, cluster of two CTAs)
};

// Whether the CTA-pair kernel can run this problem (and is worth it).
bool pair_eligible(int rows, int ncols, bool y_mn_major0, bool y_mn_major1);
// 1 = single CTAs, 2 = CTA pairs, 4 = clusters of two pairs sharing the Y slab by TMA multicast
int choose_cg(int epi, int rows, int ncols, bool y_mn_major0, bool y_mn_major1, int tile_n);

// Builds tensor maps + args for out = X0*Y0^T (+ second accumulator X1*Y1^T).
//   rows  : valid rows of X (output rows);  ncols: output columns (multiple of 32)
//   kdim0 : contraction length of phase 0;  kdim1: of phase 1 (0 = none)
std::string plan_gemm(GemmLaunch* L, const GemmOperand& X0, const GemmOperand& Y0, long long kdim0,
                      const GemmOperand* X1, const GemmOperand* Y1, long long kdim1, int rows,
                      int ncols, int splits_hint, int num_sms, int cg = 1, int tile_n = 0);

// Adds operand pair number `seg` (1 or 2) to phase 0: acc0 += Xs * Ys' over the same contraction.
// Must be called after plan_gemm and before any split bookkeeping is read.
std::string add_segment(GemmLaunch* L, int seg, const GemmOperand& X, const GemmOperand& Y, int num_sms,
                        int splits_hint);

// EPI_HUPDATE: let the kernel stage the H master tile through shared memory by TMA.
std::string set_h_prefetch(GemmLaunch* L, const float* Hm, long long n, long long Kp, long long ldh);
// EPI_RESID / EPI_KLQ: same staging for the tile of V (m x n column-major, leading dimension ldv).
std::string set_v_prefetch(GemmLaunch* L, const float* V, long long m, long long n, long long ldv);

// Column splits of a kl_fused / ab_fused launch (work items of `per` column tiles, the last split takes the
// remainder) chosen by playing the launch through a list scheduler of `slots` resident CTA pairs; see gemm_host.cu.
// per is a multiple of `chunk` and at most max_per (0 = no bound).  Returns the number of splits.
int choose_kl_splits(int pairs, int total_tiles, int rows, int Kp, int slots, int chunk, int* per_out, int max_per = 0);

// Tail helpers (GemmArgs::sk_*): how many helper CTA pairs a planned CTA-pair launch of `tiles` row tiles can
// use when `reserve_sms` SMs are to stay free, and the k-block at which the primaries hand over to them.
// Returns 0 when the launch is not eligible (single CTAs, several column chunks, split-K, segments, a grid
// that already fills the GPU or one so small that split-K is the better tool).
int plan_tail_helpers(const GemmLaunch& L, int epi, int num_sms, int reserve_sms, int* kp_out);
int balance_tail_helpers(int tiles, int helpers, int nkb0, int* kp_out);  // the pure part of the above
// Adds the helpers to the grid.  part: tiles * ncols * 256 floats; flags: 2 * tiles words, zero-initialised.
void set_tail_helpers(GemmLaunch* L, int helpers, int kp, float* part, unsigned int* flags);

std::string launch_gemm(const GemmLaunch& L, int epi, cudaStream_t stream);

}  // namespace nmfb
