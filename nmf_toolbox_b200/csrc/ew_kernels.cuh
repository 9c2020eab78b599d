// Vectorised-HBM / warp-shuffle kernels around the tensor-core contractions:
// data preparation, the element-wise halves of the multiplicative updates, the
// column / row reductions they need, the cost and stop test, the convolutive
// shift-and-fold, and Hoyer's projection.  Reductions accumulate in fp64.
//
// Device layouts (all fp32):
//   V   [n][ldv]    column-major m x n (as MATLAB supplies it), ldv % 4 == 0
//   W   [Kp][ldw]   column-major m x K;   column k is contiguous
//   H   [Kp][ldh]   ROW-major K x n;      row k is contiguous
// so "vector c of a factor" (a column of W or a row of H) is always contiguous.
// Kp = K rounded up to 32; padding vectors are all-zero and stay zero.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace nmfb {

#ifndef NMFB_EPS
#define NMFB_EPS 2.220446049250313e-16f
#endif

#define NMFB_STOP_GUARD(stop) \
  if ((stop) != nullptr && *(stop) != 0) return

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum of NV doubles per thread; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh /* [32*NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect sh from a previous use
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[warp * NV + i] = v[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < nw ? sh[lane * NV + i] : 0.0;
      v[i] = warp_sum(x);
    }
  }
}

// ---------------------------------------------------------------- V preparation
// stats[0] += sum v^2, [1] += sum v, [2] += sum v*log(v); stats_u[0] = max bits, stats_u[1] = any negative
__global__ void v_stats_kernel(const float* __restrict__ V, int m, int n, long long ldv,
                               double* stats, unsigned int* stats_u, int want_log) {
  __shared__ double sh[32 * 3];
  double acc[3] = {0.0, 0.0, 0.0};
  float mx = 0.f;
  bool neg = false;
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const float* col = V + static_cast<long long>(j) * ldv;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const float v = col[i];
      acc[0] += static_cast<double>(v) * v;
      acc[1] += v;
      if (want_log) acc[2] += static_cast<double>(v * logf(v));
      mx = fmaxf(mx, v);
      neg |= v < 0.f;
    }
  }
  block_sum<3>(acc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(stats + 0, acc[0]);
    atomicAdd(stats + 1, acc[1]);
    atomicAdd(stats + 2, acc[2]);
  }
  const unsigned int mb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(mx, 0.f)));
  const bool anyneg = __any_sync(0xffffffffu, neg);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(stats_u, mb);
    if (anyneg) atomicOr(stats_u + 1, 1u);
  }
}

// Vt = tf32_rn(V / div) (div = 1: plain rounding); sumsq += sum Vt^2.
__global__ void v_prepare_kernel(const float* __restrict__ V, float* __restrict__ Vt, int m, int n,
                                 long long ldv_in, long long ldv_out, const unsigned int* maxbits,
                                 int do_round, double* sumsq) {
  __shared__ double sh[32];
  const float div = maxbits ? __uint_as_float(*maxbits) : 1.f;
  double acc[1] = {0.0};
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const float* col = V + static_cast<long long>(j) * ldv_in;
    float* out = Vt + static_cast<long long>(j) * ldv_out;
    for (int i = threadIdx.x; i < ldv_out; i += blockDim.x) {
      float v = 0.f;
      if (i < m) {
        v = col[i];
        if (maxbits) v = v / div;
        if (do_round) v = tf32_rn(v);
      }
      out[i] = v;
      acc[0] += static_cast<double>(v) * v;
    }
  }
  block_sum<1>(acc, sh);
  if (threadIdx.x == 0 && sumsq) atomicAdd(sumsq, acc[0]);
}

// ---------------------------------------------------------------- generic helpers
// dst[r][c] (ld_dst) = src[c][r] (ld_src) for r < rows, c < cols; rows beyond / cols beyond untouched.
__global__ void transpose_kernel(const float* __restrict__ src, long long ld_src,
                                 float* __restrict__ dst, long long ld_dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int c = c0 + y, r = r0 + threadIdx.x;  // read src[c][r], r contiguous
    tile[y][threadIdx.x] = (c < cols && r < rows) ? src[static_cast<long long>(c) * ld_src + r] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int r = r0 + y, c = c0 + threadIdx.x;  // write dst[r][c], c contiguous
    if (r < rows && c < cols) dst[static_cast<long long>(r) * ld_dst + c] = tile[threadIdx.x][y];
  }
}

// dst = tf32_rn(src) over nvec vectors of length len
__global__ void round_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int nvec,
                                  int len, long long ld, const int* stop) {
  NMFB_STOP_GUARD(stop);
  for (int c = blockIdx.y; c < nvec; c += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
      dst[c * ld + i] = tf32_rn(src[c * ld + i]);
}

// hi = tf32(x), lo = tf32(x - hi): x = hi + lo to ~2^-22 relative ("3xTF32" operand split)
__global__ void split_copy_kernel(const float* __restrict__ src, float* __restrict__ hi,
                                  float* __restrict__ lo, int nvec, int len, long long ld,
                                  const int* skip = nullptr) {
  NMFB_STOP_GUARD(skip);
  for (int c = blockIdx.y; c < nvec; c += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
      const float x = src[c * ld + i];
      const float h = tf32_rn(x);
      hi[c * ld + i] = h;
      lo[c * ld + i] = tf32_rn(x - h);
    }
}

// out_sum[c] += sum_i M[c][i];  out_sq[c] += sum_i M[c][i]^2   (either may be null)
__global__ void vec_sums_kernel(const float* __restrict__ M, int nvec, int len, long long ld,
                                double* out_sum, double* out_sq, const int* stop) {
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[64];
  const int c = blockIdx.y;
  const float* v = M + static_cast<long long>(c) * ld;
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    const float x = v[i];
    acc[0] += x;
    acc[1] += static_cast<double>(x) * x;
  }
  block_sum<2>(acc, sh);
  if (threadIdx.x == 0) {
    if (out_sum) atomicAdd(out_sum + c, acc[0]);
    if (out_sq) atomicAdd(out_sq + c, acc[1]);
  }
}

// Sum split-K slabs of a Kp x Kp Gram matrix; write fp32 and tf32-rounded copies.
// gate (optional): the last block to finish publishes gate_value for a concurrently running
// consumer (ptx.cuh: gate_publish / gate_wait).
__global__ void gram_reduce_kernel(const float* __restrict__ parts, int splits, long long slab,
                                   float* __restrict__ g32, float* __restrict__ gtf,
                                   float* __restrict__ glo, int count, const int* stop,
                                   unsigned int* ticket = nullptr, unsigned int* gate = nullptr,
                                   unsigned int gate_value = 0) {
  NMFB_STOP_GUARD(stop);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (gate != nullptr) {
    if (i < count) {
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      int z = 0;
      for (; z + 4 <= splits; z += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s4[u] += parts[(z + u) * slab + i];
      }
      for (; z < splits; ++z) s4[0] += parts[z * slab + i];
      const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      g32[i] = s;
      const float hi = tf32_rn(s);
      gtf[i] = hi;
      if (glo) glo[i] = tf32_rn(s - hi);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        *ticket = 0u;
        __threadfence();
        gate_publish(gate, gate_value);
      }
    }
    return;
  }
  if (i >= count) return;
  float s4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains keep the loads in flight
  int z = 0;
  for (; z + 4 <= splits; z += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) s4[u] += parts[(z + u) * slab + i];
  }
  for (; z < splits; ++z) s4[0] += parts[z * slab + i];
  const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  g32[i] = s;
  const float hi = tf32_rn(s);
  gtf[i] = hi;
  if (glo) glo[i] = tf32_rn(s - hi);
}
// Same for a general matrix with split-K slabs (fp32 result only).
__global__ void split_reduce_kernel(const float* __restrict__ parts, int splits, long long slab,
                                    float* __restrict__ out, long long count, const int* stop) {
  NMFB_STOP_GUARD(stop);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += parts[z * slab + i];
    out[i] = s;
  }
}

// Two slab stacks in one launch (the two accumulators of ab_fused.cuh)
__global__ void split_reduce2_kernel(const float* __restrict__ parts_a, const float* __restrict__ parts_b, int splits,
                                     long long slab, float* __restrict__ out_a, float* __restrict__ out_b, long long count,
                                     const int* stop) {
  NMFB_STOP_GUARD(stop);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float sa = 0.f, sb = 0.f;
    for (int z = 0; z < splits; ++z) {
      sa += parts_a[z * slab + i];
      sb += parts_b[z * slab + i];
    }
    out_a[i] = sa;
    out_b[i] = sb;
  }
}

// Per-column coefficients of the generic W step
//   W' = W .* (A + W*p_c) ./ max(Bterm + W*q_c + lambda, eps)
// Euclidean (nmf.m:149-150): p = <W_c,B_c>, q = <W_c,A_c>, Bterm = B
// KL        (nmf.m:152-153): p = hs_c*ws_c, q = <W_c,R_c>, Bterm = hs_c
// LNMF      (lnmf.m:74-75):  W' = W .* R ./ max(hs_c, eps), then unit column SUM instead of unit L2
enum { WSTEP_EUCLID = 0, WSTEP_KL = 1, WSTEP_PLAIN = 2, WSTEP_LNMF = 3 };
// The whole W step of nmf.m:149-169 / cnmf.m:187-199 in ONE launch, one CTA per basis vector:
// everything the step needs is local to a column of W (cnmf: to the T frame-columns of one
// basis), so the three dependent reductions are block-level and the column stays in registers:
//   1  a_c = <W_c, A_c>, b_c = <W_c, B_c>                     (the diag(diag(.)) terms)
//   2  W' = W .* (A + W p_c) ./ max(Bterm + W q_c + lambda, eps)            (nmf.m:168)
//   3  W = W' / |W'_c| (nmf.m:169)  or  W(:,k,:) / (|W(:,k,:)|_F / T) (cnmf.m:196-199);
//      tf32 copy; wsum[c] = sum_i W_ic
// W, A and B are read once and W written once when T*m <= kWThreads*kWCache; longer columns
// are re-read (L2 resident).
constexpr int kWThreads = 1024;
constexpr int kWCache = 16;
struct WStepArgs {
  int mode;        // WSTEP_EUCLID / WSTEP_KL
  float* W;        // master, updated in place
  float* Wt;       // tf32 copy
  const float* A;  // numerator  (V H' or (V./V_hat) H')
  const float* B;  // denominator matrix (Euclid: W (H H')); null for KL
  int m;
  long long ld;
  int K, T;        // CTA k handles columns k + K*t, t < T (T = 1: plain nmf)
  int cnmf_style;  // 1: normalise per basis over all frames by |.|_F / T (CTA = basis, loops over frames)
                   // 2: no normalisation here: W' is stored, norm2_out[c] = |W'_c|^2 and a following
                   //    w_normalize_kernel scales (cnmf with one CTA per frame-column: T x more CTAs)
  double* norm2_out;
  double* wsum;    // [K*T] column sums of the new W (KL: holds the old ones on entry)
  const double* hs;  // KL: row sums of H
  float lambda;
  const int* stop;
  float expo;      // AB divergence: both gradients are raised to 1/alpha (dual: 1/beta) first
                   // (nmf.m:159-163); 0 or 1 = plain ratio
  const float* lambda_k;  // optional per-basis lambda_W / fixed flags (multi-source runs, nmf.m:145,168)
  const int* fixed_k;
};
// THREADS: 1024 for long columns; 256 for columns of up to 4096 rows (cnmf: hundreds of short frame-columns -
// four resident blocks per SM and four times cheaper block reductions)
template <bool CACHED, int THREADS>
__global__ void __launch_bounds__(THREADS) w_step_kernel(WStepArgs a) {
  NMFB_STOP_GUARD(a.stop);
  __shared__ double sh[32 * 2];
  __shared__ double bc[4];
  const int k = blockIdx.x;
  const int tid = threadIdx.x;
  const bool lnmf = a.mode == WSTEP_LNMF;
  const bool kl = a.mode == WSTEP_KL || lnmf;  // no B matrix
  const int total = a.T * a.m;  // elements of this basis
  if (a.fixed_k != nullptr && a.fixed_k[k] != 0) return;  // basis of a fixed source: untouched, not renormalised
  const float lambda = a.lambda_k != nullptr ? a.lambda_k[k] : a.lambda;
  float wv[kWCache], av[kWCache], bv[kWCache];
  double norm_basis = 0.0;

  for (int t = 0; t < a.T; ++t) {
    const int c = k + a.K * t;
    const long long off = static_cast<long long>(c) * a.ld;
    // ---- 1: column dots
    float s0 = 0.f, s1 = 0.f;
    if (CACHED) {
#pragma unroll
      for (int q = 0; q < kWCache; ++q) {
        const int i = tid + q * THREADS;
        const bool ok = i < a.m;
        wv[q] = ok ? a.W[off + i] : 0.f;
        av[q] = ok ? a.A[off + i] : 0.f;
        bv[q] = (ok && !kl) ? a.B[off + i] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < kWCache; ++q) {
        s0 = fmaf(wv[q], av[q], s0);
        s1 = fmaf(wv[q], bv[q], s1);
      }
    } else {
      for (int i = tid; i < a.m; i += THREADS) {
        const float w = a.W[off + i];
        s0 = fmaf(w, a.A[off + i], s0);
        if (!kl) s1 = fmaf(w, a.B[off + i], s1);
      }
    }
    double acc[2] = {s0, s1};
    block_sum<2>(acc, sh);
    if (tid == 0) {
      bc[0] = acc[0];
      bc[1] = acc[1];
    }
    __syncthreads();
    float pc, qc, bterm = 0.f;
    if (lnmf) {  // lnmf.m:74
      pc = 0.f;
      qc = 0.f;
      bterm = static_cast<float>(a.hs[c]);
    } else if (kl) {  // nmf.m:152-153
      pc = static_cast<float>(a.hs[c] * a.wsum[c]);
      qc = static_cast<float>(bc[0]);
      bterm = static_cast<float>(a.hs[c]);
    } else {   // nmf.m:149-150
      pc = static_cast<float>(bc[1]);
      qc = static_cast<float>(bc[0]);
    }
    // ---- 2: multiplicative step + column norm
    float s2 = 0.f;
    const bool powered = a.expo != 0.f && a.expo != 1.f;
    if (CACHED) {
#pragma unroll
      for (int q = 0; q < kWCache; ++q) {
        const float w = wv[q];
        float neg = av[q] + w * pc;
        float pos = (kl ? bterm : bv[q]) + w * qc;
        if (powered) {
          neg = powf(neg, a.expo);
          pos = powf(pos, a.expo);
        }
        const float wn = (tid + q * THREADS < a.m) ? w * (neg / fmaxf(pos + lambda, NMFB_EPS)) : 0.f;
        wv[q] = wn;
        s2 = lnmf ? s2 + wn : fmaf(wn, wn, s2);
      }
    } else {
      for (int i = tid; i < a.m; i += THREADS) {
        const float w = a.W[off + i];
        float neg = a.A[off + i] + w * pc;
        float pos = (kl ? bterm : a.B[off + i]) + w * qc;
        if (powered) {
          neg = powf(neg, a.expo);
          pos = powf(pos, a.expo);
        }
        const float wn = w * (neg / fmaxf(pos + lambda, NMFB_EPS));
        a.W[off + i] = wn;
        s2 = lnmf ? s2 + wn : fmaf(wn, wn, s2);
      }
    }
    acc[0] = s2;
    acc[1] = 0.0;
    block_sum<2>(acc, sh);
    if (tid == 0) bc[2] = acc[0];
    __syncthreads();
    norm_basis += bc[2];
    if (a.cnmf_style == 2) {
      if (tid == 0) a.norm2_out[c] = bc[2];
      if (CACHED) {
#pragma unroll
        for (int q = 0; q < kWCache; ++q) {
          const int i = tid + q * THREADS;
          if (i < a.m) a.W[off + i] = wv[q];
        }
      }
    } else if (!a.cnmf_style) {
      // ---- 3 (nmf): unit L2 column, tf32 copy, column sum
      const float mul = static_cast<float>(lnmf ? 1.0 / bc[2] : 1.0 / sqrt(bc[2]));  // lnmf.m:75 / nmf.m:169
      float s3 = 0.f;
      if (CACHED) {
#pragma unroll
        for (int q = 0; q < kWCache; ++q) {
          const int i = tid + q * THREADS;
          if (i < a.m) {
            const float w = wv[q] * mul;
            a.W[off + i] = w;
            a.Wt[off + i] = tf32_rn(w);
            s3 += w;
          }
        }
      } else {
        for (int i = tid; i < a.m; i += THREADS) {
          const float w = a.W[off + i] * mul;
          a.W[off + i] = w;
          a.Wt[off + i] = tf32_rn(w);
          s3 += w;
        }
      }
      acc[0] = s3;
      acc[1] = 0.0;
      block_sum<2>(acc, sh);
      if (tid == 0) a.wsum[c] = acc[0];
    } else if (CACHED) {
      // cnmf: the scale needs all T frames; park W' and come back
#pragma unroll
      for (int q = 0; q < kWCache; ++q) {
        const int i = tid + q * THREADS;
        if (i < a.m) a.W[off + i] = wv[q];
      }
    }
    __syncthreads();
  }
  if (a.cnmf_style == 1) {
    // ---- 3 (cnmf.m:196-199): W(:,k,:) /= |W(:,k,:)|_F / T
    const float div = static_cast<float>(sqrt(norm_basis) / a.T);
    for (int t = 0; t < a.T; ++t) {
      const int c = k + a.K * t;
      const long long off = static_cast<long long>(c) * a.ld;
      float s3 = 0.f;
      for (int i = tid; i < a.m; i += THREADS) {
        const float w = a.W[off + i] / div;
        a.W[off + i] = w;
        a.Wt[off + i] = tf32_rn(w);
        s3 += w;
      }
      double acc[2] = {s3, 0.0};
      block_sum<2>(acc, sh);
      if (tid == 0) a.wsum[c] = acc[0];
      __syncthreads();
    }
  }
  (void)total;
}

// Column normalisation + tf32 copy + column sums.
//   T == 0: no scaling (copy / round only)
//   T == 1 (nmf.m:133,169):    W_c *= 1/sqrt(norm2[c])
//   T >= 1, cnmf (cnmf.m:160-162,196-199): W(:,k,:) /= sqrt(sum_t norm2[k + K*t]) / T
// hscale (optional, cnmf.m:163): receives the per-basis norm so H can be compensated at init.
__global__ void w_normalize_kernel(float* __restrict__ W, float* __restrict__ Wt, int m, long long ld,
                                   int K, int T, int cnmf_style, const double* norm2, double* wsum,
                                   float* hscale, const int* stop) {
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[32];
  const int c = blockIdx.y;  // < K*max(T,1)
  const long long off = static_cast<long long>(c) * ld;
  float mul = 1.f, div = 1.f;
  if (T >= 1) {
    if (cnmf_style) {
      double s = 0.0;
      const int k = c % K;
      for (int t = 0; t < T; ++t) s += norm2[k + K * t];
      div = static_cast<float>(sqrt(s) / T);
      if (hscale && c < K && blockIdx.x == 0 && threadIdx.x == 0) hscale[c] = div;
    } else {
      mul = static_cast<float>(1.0 / sqrt(norm2[c]));
    }
    if (cnmf_style == 2) {  // lnmf.m:63: norm2 holds the column SUM
      div = 1.f;
      mul = static_cast<float>(1.0 / norm2[c]);
    }
  }
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    float w = W[off + i];
    if (T >= 1) w = (cnmf_style == 1) ? (w / div) : (w * mul);
    W[off + i] = w;
    Wt[off + i] = tf32_rn(w);
    acc[0] += w;
  }
  block_sum<1>(acc, sh);
  if (threadIdx.x == 0 && wsum) atomicAdd(wsum + c, acc[0]);
}

// H(k,:) *= scale[k] (cnmf.m:163), in place, + tf32 copy
__global__ void row_scale_kernel(float* __restrict__ H, float* __restrict__ Ht, int len, long long ld,
                                 const float* scale, int invert) {
  const int c = blockIdx.y;
  const float s = invert ? 1.f / scale[c] : scale[c];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    const float h = H[c * ld + i] * s;
    H[c * ld + i] = h;
    if (Ht) Ht[c * ld + i] = tf32_rn(h);
  }
}

// out += <GA, GB> over count fp32 entries (fp64 accumulation); feeds the trace form of the cost
__global__ void gram_dot_kernel(const float* __restrict__ GA, const float* __restrict__ GB, int count,
                                double* out, const int* stop) {
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[32];
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    acc[0] += static_cast<double>(GA[i]) * GB[i];
  block_sum<1>(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, acc[0]);
}

// ---------------------------------------------------------------- cost + stop test
struct CostArgs {
  int mode;             // 0 Euclid trace, 1 Euclid direct (scal[0] = sum (V-Vhat)^2), 2 KL, 3 sparsity terms only,
                        // 4 IS (scal[2] = divergence), 5 AB (scal[2] = bracket sum of nmf.m:214, times ab_scale)
  double ab_scale;      // -1 / (alpha * beta)
  int iter;             // 0-based index into cost[]
  int Kp;
  const float* GW;      // Kp x Kp fp32 Gram matrices (trace mode)
  const float* GH;
  double vsq;           // sum V^2                 (Euclid)
  const double* vstats; // [1] sum V, [2] sum V log V (KL)
  double* scal;         // [0] <N,H> [1] sum H [2],[3] cost-epilogue sums [4] <G_W,G_H>; reset here
  const double* wsum;   // per-column sums of W (Kp_w entries)
  int n_wsum;
  int stop_le;          // lnmf.m:88: stop if cost <= previous and the decrease <= tolerance
  const float* lamw_k;  // optional per-basis lambda_W: the W term is sum_k lamw_k[k] wsum[k] and the
                        // H-step kernels have already weighted scal[1] (host passes lambda_w = lambda_h = 1)
  double lambda_w, lambda_h;
  double tolerance;
  double* cost;
  int* stop;            // [0] stop flag, [1] number of valid cost entries
  double scale;         // multi-GPU: scal/GH are already all-reduced; nothing to do here
};
__global__ void cost_kernel(CostArgs a) {
  if (a.stop[0] != 0) return;
  __shared__ double sh[64];
  double acc[2] = {0.0, 0.0};
  acc[0] = threadIdx.x == 0 ? a.scal[4] : 0.0;
  for (int i = threadIdx.x; i < a.n_wsum; i += blockDim.x)
    acc[1] += a.lamw_k != nullptr ? a.wsum[i] * static_cast<double>(a.lamw_k[i]) : a.wsum[i];
  block_sum<2>(acc, sh);
  if (threadIdx.x != 0) return;
  double c = 0.0;
  if (a.mode == 0) {
    c = 0.5 * (a.vsq - 2.0 * a.scal[0] + acc[0]);
  } else if (a.mode == 1) {
    c = 0.5 * a.scal[2];
  } else if (a.mode == 2) {
    c = a.vstats[2] - a.scal[2] - a.vstats[1] + a.scal[3];
  } else if (a.mode == 4) {
    c = a.scal[2];
  } else if (a.mode == 5) {
    c = a.ab_scale * a.scal[2];
  }
  c += a.lambda_w * acc[1] + a.lambda_h * a.scal[1];
  a.cost[a.iter] = c;
  a.stop[1] = a.iter + 1;
  if (a.iter > 0) {
    const double prev = a.cost[a.iter - 1];
    if (a.stop_le ? (c <= prev && prev - c <= a.tolerance) : (c < prev && prev - c < a.tolerance))
      a.stop[0] = 1;  // nmf.m:221-224 / lnmf.m:88
  }
  a.scal[0] = a.scal[1] = a.scal[2] = a.scal[3] = a.scal[4] = 0.0;
}

// gram_reduce of G_H fused with <G_W, G_H> and the cost / stop test of the previous iteration:
// every block reduces its share of the split-K slabs and adds its part of the inner product;
// the last block to finish (ticket counter) finalises the cost exactly as cost_kernel does.
__global__ void gram_reduce_cost_kernel(const float* __restrict__ parts, int splits, long long slab,
                                        float* __restrict__ g32, float* __restrict__ gtf, int count,
                                        unsigned int* ticket, CostArgs c, int with_cost,
                                        unsigned int* gate = nullptr, unsigned int gate_value = 0) {
  if (c.stop[0] != 0) return;
  __shared__ double sh[64];
  __shared__ bool last;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  if (i < count) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    int z = 0;
    for (; z + 4 <= splits; z += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s4[u] += parts[(z + u) * slab + i];
    }
    for (; z < splits; ++z) s4[0] += parts[z * slab + i];
    const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    g32[i] = s;
    gtf[i] = tf32_rn(s);
    if (with_cost) acc[0] = static_cast<double>(s) * c.GW[i];
  }
  if (!with_cost && gate == nullptr) return;
  block_sum<2>(acc, sh);
  __threadfence();
  if (threadIdx.x == 0) {
    if (with_cost) atomicAdd(c.scal + 4, acc[0]);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (!with_cost) {
    if (threadIdx.x == 0) {
      *ticket = 0u;
      __threadfence();
      gate_publish(gate, gate_value);
    }
    return;
  }
  acc[0] = 0.0;
  acc[1] = 0.0;
  for (int k = threadIdx.x; k < c.n_wsum; k += blockDim.x)
    acc[1] += c.lamw_k != nullptr ? c.wsum[k] * static_cast<double>(c.lamw_k[k]) : c.wsum[k];
  block_sum<2>(acc, sh);
  if (threadIdx.x != 0) return;
  *ticket = 0u;
  volatile double* sc = c.scal;
  double cost = 0.5 * (c.vsq - 2.0 * sc[0] + sc[4]);
  cost += c.lambda_w * acc[1] + c.lambda_h * sc[1];
  c.cost[c.iter] = cost;
  c.stop[1] = c.iter + 1;
  if (c.iter > 0) {
    const double prev = c.cost[c.iter - 1];
    if (cost < prev && prev - cost < c.tolerance) c.stop[0] = 1;  // nmf.m:221-224
  }
  sc[0] = sc[1] = sc[2] = sc[3] = sc[4] = 0.0;
  if (gate != nullptr) {
    __threadfence();
    gate_publish(gate, gate_value);
  }
}

// Unfused H update for problems with too few sample tiles to fill the GPU (small column
// shards): N = W'V was formed with split-K and summed, D = (W'W) H stored next to it.
//   H <- H .* N ./ max(D + lambda, eps)  (nmf.m:180-181,199); scal[0] += <N, tf32(Hnew)>, scal[1] += sum Hnew
__global__ void h_finish_kernel(const float* __restrict__ N, const float* __restrict__ D, float* __restrict__ Hm,
                                float* __restrict__ Ht, long long ld, int n, float lambda, int freeze,
                                double* scal, const int* stop, float expo = 0.f,
                                const float* lambda_k = nullptr, const int* fixed_k = nullptr,
                                int splits = 1, long long slab = 0 /* N, D given as `splits` partial slabs */) {
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[64];
  const int k = blockIdx.y;
  double acc[2] = {0.0, 0.0};
  // per-basis settings of a multi-source run: scal[1] then receives the lambda-weighted sum
  if (lambda_k != nullptr) lambda = lambda_k[k];
  if (fixed_k != nullptr && fixed_k[k] != 0) freeze = 1;
  const double wgt = lambda_k != nullptr ? static_cast<double>(lambda) : 1.0;
  const bool powered = expo != 0.f && expo != 1.f;  // AB divergence (nmf.m:190-194)
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const long long o = static_cast<long long>(k) * ld + j;
    float nv = N[o];
    for (int z = 1; z < splits; ++z) nv += N[z * slab + o];
    float hv = Hm[o];
    if (!freeze) {
      float dv = D[o];
      for (int z = 1; z < splits; ++z) dv += D[z * slab + o];
      if (powered) {
        nv = powf(nv, expo);
        dv = powf(dv, expo);
      }
      hv = hv * (powered ? nv / fmaxf(dv + lambda, NMFB_EPS) : __fdividef(nv, fmaxf(dv + lambda, NMFB_EPS)));
      Hm[o] = hv;
    }
    const float hr = tf32_rn(hv);
    if (!freeze) Ht[o] = hr;
    acc[0] += static_cast<double>(nv) * hr;
    acc[1] += hv;
  }
  acc[1] *= wgt;
  block_sum<2>(acc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(scal + 0, acc[0]);
    atomicAdd(scal + 1, acc[1]);
  }
}

// ---------------------------------------------------------------- label-constrained factor (constrainednmf.m)
// H = Z*A with A the 0/1 label-indicator matrix of constrainednmf.m:166-170 (samples ordered so that one
// class is contiguous): column j of H is column col2z[j] of Z.
__global__ void tied_gather_kernel(const float* __restrict__ Z, long long ldz, const int* __restrict__ col2z,
                                   float* __restrict__ Hm, float* __restrict__ Ht, long long ldh, int n,
                                   const int* stop) {
  NMFB_STOP_GUARD(stop);
  const int k = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const float z = Z[k * ldz + col2z[j]];
    Hm[k * ldh + j] = z;
    Ht[k * ldh + j] = tf32_rn(z);
  }
}
// Z step (constrainednmf.m:213-237): with N = W'Qn and D = W'Qp (K x n) the gradients are N*A' and D*A',
// i.e. sums over the samples [seg[z], seg[z+1]) that share column z of Z:
//   Z <- Z .* (N A')^e ./ max((D A')^e + lambda, eps);   H = Z A
// One warp per (k, z).  N may be given as split slabs (fused KL kernel); D as a matrix or, for KL, as the
// per-basis value dvec[k] = sum_i W_ik (constrainednmf.m:219: W' * ones(m, n) * A').
// scal[0] += <N, tf32(H_new)>, scal[1] += sum(Z_new)  (the sparsity term is on Z, constrainednmf.m:251)
__global__ void tied_update_kernel(const float* __restrict__ Nparts, int splits, long long slab, long long ldn,
                                   const float* __restrict__ D, const float* __restrict__ dvec,
                                   float* __restrict__ Z, long long ldz, float* __restrict__ Hm,
                                   float* __restrict__ Ht, long long ldh, const int* __restrict__ seg, int nz, int K,
                                   float lambda, int freeze, float expo, double* scal, const int* stop) {
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[64];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long total = static_cast<long long>(K) * nz;
  const bool powered = expo != 0.f && expo != 1.f;
  double acc[2] = {0.0, 0.0};
  for (long long p = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); p < total;
       p += static_cast<long long>(gridDim.x) * warps_per_block) {
    const int k = static_cast<int>(p / nz), z = static_cast<int>(p % nz);
    const int j0 = seg[z], j1 = seg[z + 1];
    float neg = 0.f, pos = 0.f;
    for (int j = j0 + lane; j < j1; j += 32) {
      float nv = 0.f;
      for (int sp = 0; sp < splits; ++sp) nv += Nparts[sp * slab + k * ldn + j];
      neg += nv;
      if (D != nullptr) pos += D[k * ldh + j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      neg += __shfl_xor_sync(0xffffffffu, neg, o);
      pos += __shfl_xor_sync(0xffffffffu, pos, o);
    }
    if (D == nullptr) pos = dvec[k] * static_cast<float>(j1 - j0);
    const float neg_raw = neg;
    if (powered) {
      neg = powf(neg, expo);
      pos = powf(pos, expo);
    }
    float zv = Z[k * ldz + z];
    if (!freeze) zv = zv * (neg / fmaxf(pos + lambda, NMFB_EPS));  // constrainednmf.m:235
    const float zt = tf32_rn(zv);
    if (!freeze) {
      if (lane == 0) Z[k * ldz + z] = zv;
      for (int j = j0 + lane; j < j1; j += 32) {  // H = Z*A (constrainednmf.m:237)
        Hm[k * ldh + j] = zv;
        Ht[k * ldh + j] = zt;
      }
    }
    if (lane == 0) {
      acc[0] += static_cast<double>(neg_raw) * zt;
      acc[1] += zv;
    }
  }
  block_sum<2>(acc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(scal + 0, acc[0]);
    atomicAdd(scal + 1, acc[1]);
  }
}

// ---------------------------------------------------------------- convolutive helpers
// Hs[k + K*t][j] = tf32(H[k][j - t]) for j >= t, else 0   (cnmf.m:188, RFD.m:37)
// Column shards (several GPUs): H points at the shard's first own column inside a buffer that holds `left`
// columns of the left neighbour before it (the halo), so the shift reaches back to j - t >= -left; n then
// also covers the right halo columns.  ldh / lds: leading dimensions of H and Hs.
__global__ void hstack_kernel(const float* __restrict__ H, float* __restrict__ Hs, int K, int T, int n,
                              long long ldh, long long lds, int left, const int* stop) {
  NMFB_STOP_GUARD(stop);
  const int c = blockIdx.y;  // 0 .. K*T-1
  const int k = c % K, t = c / K;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    Hs[c * lds + j] = (j + left >= t) ? tf32_rn(H[k * ldh + (j - t)]) : 0.f;
}

// neg[k][j] = sum_t P[k+K*t][j+t], pos likewise from D (cnmf.m:218-227, euclidean);
// H <- H .* neg ./ max(pos + lambda, eps) (cnmf.m:231); scal[0] += <neg, tf32(Hnew)>, scal[1] += sum Hnew
// n_src >= n: columns available in P and D (a column shard also holds the T-1 columns that follow it);
// ldp / ld: leading dimensions of P, D and of H.
__global__ void fold_update_kernel(const float* __restrict__ P, const float* __restrict__ D,
                                   float* __restrict__ H, int K, int T, int n, int n_src, long long ldp, long long ld,
                                   float lambda, int freeze, double* scal, const int* stop,
                                   float expo = 0.f, int pos_unshifted = 0, float* __restrict__ Hs_out = nullptr,
                                   long long lds = 0) {
  // Hs_out (single GPU): the new H goes straight into the shifted stack of the next iteration,
  // Hs[k + K*t][j + t] = tf32(H_new[k][j]) (cnmf.m:188), which saves the separate stacking pass
  // expo: outer exponent of both gradients (AB divergence, cnmf.m:229-232); pos_unshifted: the KL
  // branch of cnmf.m:221-222 does not shift V_pos
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[64];
  const int k = blockIdx.y;
  double acc[2] = {0.0, 0.0};
  const bool powered = expo != 0.f && expo != 1.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    float neg = 0.f, pos = 0.f;
    // branch-free body (clamped index + select) so that the 2T loads of a thread are issued together
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
      const bool ok = j + t < n_src;
      const long long row = static_cast<long long>(k + K * t) * ldp;
      const float p = P[row + (ok ? j + t : j)];
      const float d = D[row + ((ok && !pos_unshifted) ? j + t : j)];
      neg += ok ? p : 0.f;
      pos += (ok || pos_unshifted) ? d : 0.f;
    }
    if (powered) {
      neg = powf(neg, expo);
      pos = powf(pos, expo);
    }
    float h = H[k * ld + j];
    if (!freeze) {
      h = h * (neg / fmaxf(pos + lambda, NMFB_EPS));
      H[k * ld + j] = h;
    }
    const float ht = tf32_rn(h);
    if (Hs_out != nullptr && !freeze) {
      for (int t = 0; t < T; ++t)
        if (j + t < n) Hs_out[static_cast<long long>(k + K * t) * lds + j + t] = ht;
    }
    acc[0] += static_cast<double>(neg) * ht;
    acc[1] += h;
  }
  block_sum<2>(acc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(scal + 0, acc[0]);
    atomicAdd(scal + 1, acc[1]);
  }
}

// ---------------------------------------------------------------- nmfsc helpers
// Xnew = X - step * (Dp - Dn)   (nmfsc.m:148,154 / 200,205)
__global__ void grad_step_kernel(const float* __restrict__ X, const float* __restrict__ Dp,
                                 const float* __restrict__ Dn, float* __restrict__ Xnew, int nvec,
                                 int len, long long ld, double step) {
  const float s = static_cast<float>(step);
  const int c = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    const long long o = static_cast<long long>(c) * ld + i;
    Xnew[o] = X[o] - s * (Dp[o] - Dn[o]);
  }
}
// multiplicative step + tf32 head / tail of the result in one pass (nmfsc.m:232 followed by the operand split)
__global__ void mu_step_split_kernel(float* __restrict__ X, const float* __restrict__ N, const float* __restrict__ D,
                                     float* __restrict__ hi, float* __restrict__ lo, int len, long long ld,
                                     const int* skip) {
  NMFB_STOP_GUARD(skip);
  const int c = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    const long long o = static_cast<long long>(c) * ld + i;
    const float x = X[o] * (N[o] / fmaxf(D[o], NMFB_EPS));
    X[o] = x;
    const float h = tf32_rn(x);
    hi[o] = h;
    lo[o] = tf32_rn(x - h);
  }
}
// plain multiplicative step X <- X .* N ./ max(D, eps)   (nmfsc.m:182,232)
__global__ void mu_step_kernel(float* __restrict__ X, const float* __restrict__ N,
                               const float* __restrict__ D, int len, long long ld, const int* skip = nullptr) {
  NMFB_STOP_GUARD(skip);
  const int c = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    const long long o = static_cast<long long>(c) * ld + i;
    X[o] = X[o] * (N[o] / fmaxf(D[o], NMFB_EPS));
  }
}
// nmfsc.m:185-187: H rows -> unit L2, W columns scaled by the norms.  sq[c] = sum H_c^2.
__global__ void renorm_pair_kernel(float* __restrict__ H, int n, long long ldh, float* __restrict__ W,
                                   int m, long long ldw, const double* sq, const int* skip = nullptr) {
  NMFB_STOP_GUARD(skip);
  const int c = blockIdx.y;
  const float nrm = static_cast<float>(sqrt(sq[c]));
  const float inv = 1.f / nrm;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    H[c * ldh + i] = inv * H[c * ldh + i];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    W[c * ldw + i] = W[c * ldw + i] * nrm;
}

// Hoyer's projection (projfunc.m:13-55, k2 and nn as given), one CTA per vector,
// vector kept in global memory (L1/L2 resident), zero-set kept as per-thread bit masks.
constexpr int kProjThreads = 512;
constexpr int kProjMaskWords = 8;  // supports len <= 512 * 32 * 8 = 131072
// Optional fusion of what surrounds the projection in a line-search trial of nmfsc.m / cnmfsc.m:
//   before: the projected-gradient step X = src - step * (Dp - Dn) (nmfsc.m:154 / 205), formed in the
//           first sweep with the step size read from device memory (the line search lives on the device);
//   after : the tf32 head / tail split of the projected vector (operands of the split-tf32 contractions).
struct ProjFuse {
  const float* src = nullptr;
  const float* Dp = nullptr;
  const float* Dn = nullptr;
  const double* step = nullptr;
  float* hi = nullptr;
  float* lo = nullptr;
  const int* skip = nullptr;  // device flag: non-zero = this launch is a no-op
};
// EPT > 0: the vector (len <= kProjThreads * EPT) lives in registers between the fused first and last
// sweep - a pass then costs two block reductions and no memory traffic; EPT == 0: the vector stays in
// global memory (L1/L2 resident), any length up to kProjThreads * 32 * kProjMaskWords.
template <int EPT>
__global__ void __launch_bounds__(kProjThreads)
projfunc_kernel_t(float* __restrict__ X, int len, long long ld, double k1, double k2, int nn,
                  int* iters_out, int* fail_flag, ProjFuse f) {
  NMFB_STOP_GUARD(f.skip);
  __shared__ double sh[32 * 3];
  __shared__ double bc[4];
  const long long voff = static_cast<long long>(blockIdx.x) * ld;
  float* v = X + voff;
  const int tid = threadIdx.x;
  const float stepf = f.step != nullptr ? static_cast<float>(*f.step) : 0.f;
  constexpr int kWords = EPT > 0 ? (EPT + 31) / 32 : kProjMaskWords;
  uint32_t zmask[kWords];
  uint32_t negmask[kWords];  // signs when nn == 0
  float r[EPT > 0 ? EPT : 1];
#pragma unroll
  for (int w = 0; w < kWords; ++w) zmask[w] = negmask[w] = 0u;
// q-th element of this thread: e = tid + q * kProjThreads (fully unrolled when the vector is in registers)
#define NMFB_PROJ_FOR(q, e) \
  _Pragma("unroll") for (int q = 0, e = tid; (EPT > 0 ? q < EPT : e < len); ++q, e += kProjThreads) if (EPT == 0 || e < len)
#define NMFB_PROJ_GET(q, e) (EPT > 0 ? r[EPT > 0 ? q : 0] : v[e])
#define NMFB_PROJ_SET(q, e, x)       \
  do {                               \
    if (EPT > 0) r[EPT > 0 ? q : 0] = (x); \
    else v[e] = (x);                 \
  } while (0)

  // projfunc.m:16-22: v = s + (k1 - sum(s)) / N
  double acc[3] = {0.0, 0.0, 0.0};
  NMFB_PROJ_FOR(q, e) {
    float s = f.src != nullptr ? f.src[voff + e] - stepf * (f.Dp[voff + e] - f.Dn[voff + e]) : v[e];
    if (!nn && s < 0.f) {
      negmask[q >> 5] |= 1u << (q & 31);
      s = -s;
    }
    NMFB_PROJ_SET(q, e, s);
    acc[0] += s;
  }
  block_sum<3>(acc, sh);
  if (tid == 0) bc[0] = (k1 - acc[0]) / len;
  __syncthreads();
  double shift = bc[0];  // pending additive constant on all non-zeroed entries
  int nz = 0;
  int iters = 0;
  for (int pass = 0; pass < 100000; ++pass) {
    // sweep 1: apply pending shift, then a, b, c of projfunc.m:31-36
    const double mid = k1 / (len - nz);
    acc[0] = acc[1] = acc[2] = 0.0;
    NMFB_PROJ_FOR(q, e) {
      const bool z = (zmask[q >> 5] >> (q & 31)) & 1u;
      float x = 0.f;
      if (!z) {
        x = static_cast<float>(static_cast<double>(NMFB_PROJ_GET(q, e)) + shift);
        NMFB_PROJ_SET(q, e, x);
      }
      const double w = z ? 0.0 : static_cast<double>(x) - mid;
      acc[0] += w * w;
      acc[1] += w * x;
      acc[2] += static_cast<double>(x) * x;
    }
    block_sum<3>(acc, sh);
    if (tid == 0) {
      const double a = acc[0], b = 2.0 * acc[1], c = acc[2] - k2;
      const double disc = b * b - 4.0 * a * c;
      bc[0] = (-b + (disc > 0.0 ? sqrt(disc) : 0.0)) / (2.0 * a);  // real(sqrt(.)), projfunc.m:37
    }
    __syncthreads();
    const double alphap = bc[0];
    // sweep 2: v = alphap*w + v (projfunc.m:38); all(v >= 0)?; zero negatives, tempsum (49-51)
    acc[0] = acc[1] = acc[2] = 0.0;  // [0] any negative, [1] zero count, [2] tempsum
    NMFB_PROJ_FOR(q, e) {
      const bool z = (zmask[q >> 5] >> (q & 31)) & 1u;
      if (z) {
        acc[1] += 1.0;
      } else {
        const float x0 = NMFB_PROJ_GET(q, e);
        const float x = static_cast<float>(alphap * (static_cast<double>(x0) - mid) + x0);
        if (!(x >= 0.f)) acc[0] += 1.0;  // also catches NaN
        NMFB_PROJ_SET(q, e, x);          // (entries <= 0 are finalised below only if the loop continues)
        if (x <= 0.f) {
          zmask[q >> 5] |= 1u << (q & 31);
          acc[1] += 1.0;
        } else {
          acc[2] += x;
        }
      }
    }
    block_sum<3>(acc, sh);
    if (tid == 0) {
      bc[1] = acc[0];
      bc[2] = acc[1];
      bc[3] = acc[2];
    }
    __syncthreads();
    iters = pass + 1;
    if (bc[1] == 0.0) break;  // projfunc.m:40-44 (entries equal to 0 stay as they are)
    nz = static_cast<int>(bc[2]);
    if (!(bc[3] == bc[3]) || nz >= len) {  // NaN or everything zeroed: MATLAB would never return
      if (tid == 0 && fail_flag) *fail_flag = 1;
      break;
    }
    // projfunc.m:50-53: zero the set, spread (k1 - tempsum) over the rest (applied lazily in sweep 1)
    NMFB_PROJ_FOR(q, e) {
      if ((zmask[q >> 5] >> (q & 31)) & 1u) NMFB_PROJ_SET(q, e, 0.f);
    }
    shift = (k1 - bc[3]) / (len - nz);
    __syncthreads();
  }
  // last sweep: signs back (projfunc.m:58-60), result to memory, optional tf32 head / tail
  NMFB_PROJ_FOR(q, e) {
    float x = NMFB_PROJ_GET(q, e);
    if (!nn && ((negmask[q >> 5] >> (q & 31)) & 1u)) x = -x;
    v[e] = x;
    if (f.hi != nullptr) {
      const float hi = tf32_rn(x);
      f.hi[voff + e] = hi;
      f.lo[voff + e] = tf32_rn(x - hi);
    }
  }
  if (tid == 0 && iters_out) iters_out[blockIdx.x] = iters;
#undef NMFB_PROJ_FOR
#undef NMFB_PROJ_GET
#undef NMFB_PROJ_SET
}
// host-side dispatch on the vector length
inline void launch_projfunc(cudaStream_t stream, int count, float* X, int len, long long ld, double k1, double k2, int nn,
                            int* iters_out, int* fail_flag, const ProjFuse& f = ProjFuse()) {
  if (len <= kProjThreads * 8)
    projfunc_kernel_t<8><<<count, kProjThreads, 0, stream>>>(X, len, ld, k1, k2, nn, iters_out, fail_flag, f);
  else if (len <= kProjThreads * 32)
    projfunc_kernel_t<32><<<count, kProjThreads, 0, stream>>>(X, len, ld, k1, k2, nn, iters_out, fail_flag, f);
  else
    projfunc_kernel_t<0><<<count, kProjThreads, 0, stream>>>(X, len, ld, k1, k2, nn, iters_out, fail_flag, f);
}

// ---------------------------------------------------------------- device-side line search (nmfsc.m:146-179,196-229)
// The accept / halve decisions of the projected-gradient line searches are taken on the device.  The
// host queues a fixed PATTERN of kernels per iteration, every kernel guarded by the skip word of its
// phase, and the small kernels below move the phase on:
//   phase 0  H gradient (G_W, N = W'V, D = G_W H)           -> 1 (H_sparsity > 0) or 2 (multiplicative H step)
//   phase 1  one H trial: step + projfunc + split, objective  -> stays (halved) or 2 (accepted)
//   phase 2  commit H; W gradient (G_H, A = VH', B = W G_H)  -> 3 (W_sparsity > 0) or 4 (multiplicative W step)
//   phase 3  one W trial                                      -> stays or 4
//   phase 4  commit W; cost(iter+1) and stop test            -> 0, or everything off (done)
// A pattern holds a fixed number of trial slots; a search that needs more simply continues in the
// trial slots of the next pattern (all other phases of that pattern are no-ops), so the host never
// waits for a decision and the sequence of accepted / halved steps is exactly the reference's.
struct LsState {
  double stepH, stepW;  // nmfsc.m:133-134
  double begobj;        // objective the running line search must not exceed (nmfsc.m:149,197)
  int skip[5];          // per phase: 1 = kernels of that phase do nothing
  int iter;             // completed iterations
  int ncost;            // valid entries of cost[]
  int done;             // loop over: converged, step-size underflow, maxiter reached or failure
  int failed;           // projfunc produced non-finite values
  int trials;           // line-search trials evaluated so far (diagnostic)
  int maxiter;
  double tolerance;
  int* halvings;        // [2 * maxiter]: rejected trials of the H / W search of every iteration (diagnostic)
};
enum { LS_INIT = 0, LS_TO_HTRIAL, LS_H_TO_WGRAD, LS_TO_WTRIAL, LS_W_TO_COST };
// single-thread transitions that do not decide anything
__device__ inline void ls_advance_body(LsState* st, int what, volatile double* scal, double* cost, const int* fail) {
  if (what == LS_INIT) {  // nmfsc.m:138-139: cost(1) = objective of the initial factors
    cost[0] = 0.5 * scal[0];
    scal[0] = scal[1] = 0.0;
    st->ncost = 1;
    if (fail != nullptr && *fail != 0) {
      st->failed = st->done = 1;
      for (int g = 0; g < 5; ++g) st->skip[g] = 1;
    }
    return;
  }
  if (st->done) return;
  if (what == LS_TO_HTRIAL) {  // phase 0 -> 1
    if (st->skip[0]) return;
    st->begobj = cost[st->iter];  // nmfsc.m:149
    st->skip[0] = 1;
    st->skip[1] = 0;
  } else if (what == LS_H_TO_WGRAD) {  // phase 0 -> 2 (no H line search)
    if (st->skip[0]) return;
    st->skip[0] = 1;
    st->skip[2] = 0;
  } else if (what == LS_TO_WTRIAL) {  // phase 2 -> 3: begobj = objective with the new H (nmfsc.m:193,197)
    if (st->skip[2]) return;
    st->begobj = 0.5 * scal[0];
    scal[0] = scal[1] = 0.0;
    st->skip[2] = 1;
    st->skip[3] = 0;
  } else if (what == LS_W_TO_COST) {  // phase 2 -> 4 (no W line search)
    if (st->skip[2]) return;
    st->skip[2] = 1;
    st->skip[4] = 0;
  }
}
__global__ void ls_advance_kernel(LsState* st, int what, double* scal, double* cost, const int* fail) {
  ls_advance_body(st, what, scal, cost, fail);
}
// end of a trial slot: accept (nmfsc.m:164-166,178 / 215-217,228) or halve (169-174 / 220-225)
__device__ inline void ls_decide_body(LsState* st, int for_w, volatile double* scal, const int* fail) {
  const int ph = for_w ? 3 : 1;
  if (st->done || st->skip[ph]) return;
  const double newobj = 0.5 * scal[0];
  scal[0] = scal[1] = 0.0;
  ++st->trials;
  double& step = for_w ? st->stepW : st->stepH;
  if (fail != nullptr && *fail != 0) {
    st->failed = st->done = 1;
  } else if (newobj <= st->begobj) {
    step *= 1.2;
    st->skip[ph] = 1;
    st->skip[ph + 1] = 0;
    return;
  } else {
    step /= 2;
    if (st->halvings != nullptr && st->iter < st->maxiter) ++st->halvings[2 * st->iter + (for_w ? 1 : 0)];
    if (!(step < 1e-200)) return;
    st->ncost = st->iter + 1;  // 'Algorithm converged': cost = cost(1:iter) (nmfsc.m:170-174,221-225)
    st->done = 1;
  }
  for (int g = 0; g < 5; ++g) st->skip[g] = 1;
}
__global__ void ls_decide_kernel(LsState* st, int for_w, double* scal, const int* fail) {
  ls_decide_body(st, for_w, scal, fail);
}
// end of an iteration: cost(iter+1) and the stop test (nmfsc.m:237-244)
__device__ inline void ls_cost_body(LsState* st, volatile double* scal, double* cost) {
  if (st->done || st->skip[4]) return;
  const int it = st->iter + 1;  // 1-based iteration that just finished
  const double c = 0.5 * scal[0];
  scal[0] = scal[1] = 0.0;
  cost[it] = c;
  st->iter = it;
  st->ncost = it + 1;
  st->skip[4] = 1;
  const bool conv = it > 1 && c < cost[it - 1] && cost[it - 1] - c < st->tolerance;  // nmfsc.m:241-244
  if (conv || it >= st->maxiter) {
    st->done = 1;
    for (int g = 0; g < 5; ++g) st->skip[g] = 1;
  } else {
    st->skip[0] = 0;
  }
}
__global__ void ls_cost_kernel(LsState* st, double* scal, double* cost) { ls_cost_body(st, scal, cost); }
// What the LAST block of an objective kernel does with the finished sum (saves one launch per objective):
enum { LSFIN_NONE = 0, LSFIN_DECIDE_H, LSFIN_DECIDE_W, LSFIN_COST, LSFIN_TO_WTRIAL, LSFIN_INIT };
struct LsFin {
  int mode = LSFIN_NONE;
  LsState* st = nullptr;
  double* cost = nullptr;
  const int* fail = nullptr;
  unsigned int* ticket = nullptr;  // zero on entry; reset by the last block
};
__device__ inline void ls_finish(const LsFin& f, double* scal) {
  switch (f.mode) {
    case LSFIN_DECIDE_H: ls_decide_body(f.st, 0, scal, f.fail); break;
    case LSFIN_DECIDE_W: ls_decide_body(f.st, 1, scal, f.fail); break;
    case LSFIN_COST: ls_cost_body(f.st, scal, f.cost); break;
    case LSFIN_TO_WTRIAL: ls_advance_body(f.st, LS_TO_WTRIAL, scal, f.cost, f.fail); break;
    case LSFIN_INIT: ls_advance_body(f.st, LS_INIT, scal, f.cost, f.fail); break;
    default: break;
  }
}
// accepted trial -> current factor (master, tf32 head, tf32 tail) in one launch
__global__ void copy3_kernel(const float* __restrict__ a0, float* __restrict__ b0, const float* __restrict__ a1,
                             float* __restrict__ b1, const float* __restrict__ a2, float* __restrict__ b2,
                             long long count4, const int* skip) {
  NMFB_STOP_GUARD(skip);
  const float4* s0 = reinterpret_cast<const float4*>(a0);
  const float4* s1 = reinterpret_cast<const float4*>(a1);
  const float4* s2 = reinterpret_cast<const float4*>(a2);
  float4* d0 = reinterpret_cast<float4*>(b0);
  float4* d1 = reinterpret_cast<float4*>(b1);
  float4* d2 = reinterpret_cast<float4*>(b2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    d0[i] = s0[i];
    d1[i] = s1[i];
    d2[i] = s2[i];
  }
}

}  // namespace nmfb
