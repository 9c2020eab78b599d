// Row-sharded W step for column-sharded (multi-GPU) nmf runs.
//
// With V and H partitioned by columns every rank holds a PARTIAL numerator A_r = V_r H_r' (m x K) of
// the W update (nmf.m:149-153).  Instead of all-reducing the m x K partials and then repeating the
// identical W step on every rank, rank r owns a block of ROWS of W:
//
//   1  it reads its row block of every rank's partial A straight from peer memory (NVLink) and sums it
//      in rank order (the reduce-scatter half of an all-reduce, fused into the consumer),
//   2  forms the partial column dots <W_k, A_k>, <W_k, B_k> of nmf.m's diag(diag(.)) terms over its
//      rows and exchanges them with the same block of the other ranks (16 bytes per rank and column),
//   3  takes the multiplicative step (nmf.m:168) on its rows, exchanges the partial column norms and
//      sums, normalises (nmf.m:169), and
//   4  writes the tf32 operand copy of its rows into EVERY rank's W (the all-gather half), keeping the
//      fp32 master rows local (they are gathered once, at the end of the run).
//
// One launch per iteration: block b handles columns b, b + G, ... and only ever synchronises with block
// b of the other ranks (p2p_block_barrier: flags in the peers' regions, monotone epochs), so blocks of
// one GPU never wait for each other and a grid of at most one block per SM cannot deadlock.  The
// element-wise work, the small B = W G_H product and the traffic of the second half of the all-reduce
// shrink with the number of ranks; what remains replicated is O(K^2).
#pragma once
#include <algorithm>

#include "comm.cuh"
#include "ew_kernels.cuh"

namespace nmfb {

constexpr int kWsThreads = 512;
constexpr int kWsCache = 4;  // float4 per thread: row blocks of up to 4 * 4 * 512 = 8192 rows

struct WShardArgs {
  PeerTable t;
  int mode;           // WSTEP_EUCLID (also IS / AB with expo) or WSTEP_KL
  int K;              // basis columns
  int r0, mb;         // this rank's rows [r0, r0 + mb), both multiples of 4 (padding rows included: they stay 0)
  long long ld;       // leading dimension of W, A, B
  size_t a_off;       // byte offset of the partial A (K x ld) in every rank's region
  size_t b_off;       // byte offset of a partial B in the region (IS / AB: both gradients are partial), 0 = none
  const float* Bloc;  // local B = W G_H on this rank's rows (Euclidean), or null
  float* Wm;          // fp32 master of W (local; only this rank's rows are kept current)
  size_t wt_off;      // byte offset of the tf32 operand copy of W in every rank's region
  size_t x_off;       // byte offset of the exchange slots [2][kMaxBlocks][kMaxRanks] x 32 bytes (ws_exchange)
  double* wsum;       // [K] column sums of W (KL: the old sums on entry), replicated
  const double* hs;   // KL: row sums of H, already summed over the ranks
  float lambda;
  const float* lambda_k;
  const int* fixed_k;
  float expo;         // AB: outer exponent of both gradients (nmf.m:159-163); 0 or 1 = none
  const int* stop;
  int epoch0, rounds;
  int open_barrier;   // meet the peers before touching their partial sums (needed when no collective on this
                      // stream separates the kernel from the GEMMs that produced them)
  unsigned long long* timing;  // optional [grid][8] %globaltimer stamps of the last launch (NMFB_WS_TIMING=1)
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define NMFB_WS_STAMP(i) \
  if (a.timing != nullptr && threadIdx.x == 0) a.timing[blockIdx.x * 8 + (i)] = global_ns()

__device__ __forceinline__ float4 ld_peer4(const char* base, size_t off_bytes, long long idx4) {
  return __ldcv(reinterpret_cast<const float4*>(base + off_bytes) + idx4);
}
__device__ __forceinline__ void add4(float4& s, const float4& x) {
  s.x += x.x;
  s.y += x.y;
  s.z += x.z;
  s.w += x.w;
}
// Sum over the ranks (fixed order: every rank forms the same sums) of float4 element idx4 of the buffer at
// byte offset `off` of every region.  All N loads are issued before the first add: a remote load takes
// a microsecond or two over NVLink, so the number in flight is what counts.
template <int N>
__device__ __forceinline__ float4 sum_peers4(const PeerTable& t, size_t off, long long idx4) {
  float4 x[N];
#pragma unroll
  for (int r = 0; r < N; ++r) x[r] = ld_peer4(t.base[r], off, idx4);
  float4 s = x[0];
#pragma unroll
  for (int r = 1; r < N; ++r) add4(s, x[r]);
  return s;
}

// All ranks' blocks b swap two partial sums.  NCCL's "LL" idea: the data travels WITH its flag - every 8-byte
// word carries 32 bits of payload and the 32-bit tag of this exchange (8-byte stores are single-copy atomic), so
// the receiver polls the payload words themselves and no fence or separate flag round trip is needed: one
// NVLink one-way latency per exchange.  (A slot is rewritten one exchange later at the earliest, and a rank can
// only get that far after every peer has answered the exchange in between, i.e. has consumed this one.)
// v0, v1 are taken from thread 0; the sums over the ranks (fixed order) are returned to every thread.
template <int N>
__device__ __forceinline__ void ws_exchange(const PeerTable& t, size_t site_off, uint32_t tag, double v0, double v1,
                                            double* xs /* shared, [2 * kMaxRanks + 2] */, double& out0, double& out1) {
  if (threadIdx.x == 0) {
    xs[2 * kMaxRanks] = v0;
    xs[2 * kMaxRanks + 1] = v1;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int q = threadIdx.x;
    const unsigned long long b0 = static_cast<unsigned long long>(__double_as_longlong(xs[2 * kMaxRanks]));
    const unsigned long long b1 = static_cast<unsigned long long>(__double_as_longlong(xs[2 * kMaxRanks + 1]));
    const unsigned long long tg = static_cast<unsigned long long>(tag) << 32;
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(t.base[q] + site_off) +
                                       (static_cast<size_t>(blockIdx.x) * kMaxRanks + t.rank) * 4;
    dst[0] = tg | (b0 & 0xffffffffull);
    dst[1] = tg | (b0 >> 32);
    dst[2] = tg | (b1 & 0xffffffffull);
    dst[3] = tg | (b1 >> 32);
    const volatile unsigned long long* src = reinterpret_cast<const volatile unsigned long long*>(t.base[t.rank] + site_off) +
                                             (static_cast<size_t>(blockIdx.x) * kMaxRanks + q) * 4;
    unsigned long long w[4];
    const long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      while (((w[i] = src[i]) >> 32) != tag) {
        if (clock64() - t0 > 30000000000LL) {
          printf("nmfb: peer exchange timeout (rank %d block %d waiting for rank %d, tag %u)\n", t.rank, blockIdx.x, q, tag);
          __trap();
        }
      }
    }
    xs[2 * q] = __longlong_as_double(static_cast<long long>((w[0] & 0xffffffffull) | (w[1] << 32)));
    xs[2 * q + 1] = __longlong_as_double(static_cast<long long>((w[2] & 0xffffffffull) | (w[3] << 32)));
  }
  __syncthreads();
  out0 = 0.0;
  out1 = 0.0;
#pragma unroll
  for (int r = 0; r < N; ++r) {
    out0 += xs[2 * r];
    out1 += xs[2 * r + 1];
  }
  __syncthreads();  // xs may be reused by the next exchange
}

// N = number of ranks (compile time: the peer loads are unrolled).  Only the summed numerator rows stay in
// registers between the phases (they came over NVLink); W and the local B are re-read from L2, which keeps
// the kernel at two resident blocks per SM - a grid of K <= 256 columns then needs a single round.
// BULK: the peers' rows arrive in shared memory by N TMA bulk copies per column (one request of mb*4 bytes
// per rank instead of thousands of 16-byte loads) and the finished tf32 rows leave by N bulk stores; needs
// N * mb * 4 bytes of dynamic shared memory (= the bytes of one column of W).
template <int N, bool BULK>
__global__ void __launch_bounds__(kWsThreads, 2) w_step_sharded_kernel(WShardArgs a) {
  NMFB_STOP_GUARD(a.stop);
  extern __shared__ __align__(128) uint8_t ws_smem[];
  __shared__ uint64_t ld_bar;
  __shared__ double sh[32 * 2];
  __shared__ double xs[2 * kMaxRanks + 2];
  const int tid = threadIdx.x;
  float* stage = reinterpret_cast<float*>(ws_smem);  // [N][mb] partial rows, later [mb] finished tf32 rows
  uint32_t ld_phase = 0;                              // completed uses of ld_bar (uniform over the block)
  if (BULK) {
    if (tid == 0) {
      mbar_init(&ld_bar, 1);
      fence_barrier_init();
    }
    __syncthreads();
  }
  const bool kl = a.mode == WSTEP_KL;
  const bool powered = a.expo != 0.f && a.expo != 1.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.open_barrier) p2p_block_barrier(a.t, 6 * kFlagBytes, a.epoch0);  // every rank's numerator partial is complete
  NMFB_WS_STAMP(0);
  for (int round = 0; round < a.rounds; ++round) {
    const int k = blockIdx.x + round * gridDim.x;
    const bool active = k < a.K && !(a.fixed_k != nullptr && a.fixed_k[k] != 0);
    const long long col4 = (static_cast<long long>(k) * a.ld + a.r0) >> 2;  // float4 index of (r0, k)
    const float4* Wcol = reinterpret_cast<const float4*>(a.Wm) + col4;
    const float4* Bcol = a.Bloc != nullptr ? reinterpret_cast<const float4*>(a.Bloc) + col4 : nullptr;
    float4 av[kWsCache], bv[kWsCache];  // bv only carries data when B is partial as well (IS / AB)
    float s0 = 0.f, s1 = 0.f;
    if (BULK && active) {
      // ---- 1 (bulk): every rank's rows of column k -> shared memory, one TMA request per rank
      const uint32_t bytes = static_cast<uint32_t>(a.mb) * 4u;
      if (tid == 0 && bytes > 0) {
        mbar_arrive_expect_tx(&ld_bar, bytes * N);
#pragma unroll
        for (int r = 0; r < N; ++r)
          bulk_load_1d(smem_u32(stage + static_cast<size_t>(r) * a.mb),
                       reinterpret_cast<const float4*>(a.t.base[r] + a.a_off) + col4, bytes, &ld_bar);
      }
      if (bytes > 0) {
        mbar_wait(&ld_bar, ld_phase & 1);
        ++ld_phase;
      }
#pragma unroll
      for (int q = 0; q < kWsCache; ++q) {
        const int i4 = tid + q * kWsThreads;
        float4 sacc = zero4;
        if (4 * i4 < a.mb) {
          sacc = reinterpret_cast<const float4*>(stage)[i4];
#pragma unroll
          for (int r = 1; r < N; ++r) add4(sacc, reinterpret_cast<const float4*>(stage + static_cast<size_t>(r) * a.mb)[i4]);
        }
        av[q] = sacc;
        bv[q] = zero4;
      }
    } else if (active) {
      // ---- 1: the summed numerator (and, for IS / AB, denominator) partials of this rank's rows
#pragma unroll
      for (int q = 0; q < kWsCache; ++q) {
        const int i4 = tid + q * kWsThreads;
        const bool ok = 4 * i4 < a.mb;
        av[q] = ok ? sum_peers4<N>(a.t, a.a_off, col4 + i4) : zero4;
        bv[q] = (ok && a.b_off != 0) ? sum_peers4<N>(a.t, a.b_off, col4 + i4) : zero4;
      }
    }
    if (active) {
#pragma unroll
      for (int q = 0; q < kWsCache; ++q) {
        const int i4 = tid + q * kWsThreads;
        if (4 * i4 < a.mb) {
          const float4 w = Wcol[i4];
          const float4 b = Bcol != nullptr ? Bcol[i4] : bv[q];
          s0 = fmaf(w.x, av[q].x, fmaf(w.y, av[q].y, fmaf(w.z, av[q].z, fmaf(w.w, av[q].w, s0))));
          s1 = fmaf(w.x, b.x, fmaf(w.y, b.y, fmaf(w.z, b.z, fmaf(w.w, b.w, s1))));
        }
      }
    }
    // ---- 2: partial column dots <-> every rank
    double acc[2] = {s0, s1};
    block_sum<2>(acc, sh);
    NMFB_WS_STAMP(1);
    double d0, d1;
    ws_exchange<N>(a.t, a.x_off, static_cast<uint32_t>(a.epoch0 + round), acc[0], acc[1], xs, d0, d1);
    NMFB_WS_STAMP(2);
    float pc = 0.f, qc = 0.f, bterm = 0.f, lambda = a.lambda;
    if (active) {
      if (a.lambda_k != nullptr) lambda = a.lambda_k[k];
      if (kl) {  // nmf.m:152-153
        pc = static_cast<float>(a.hs[k] * a.wsum[k]);
        qc = static_cast<float>(d0);
        bterm = static_cast<float>(a.hs[k]);
      } else {  // nmf.m:149-150
        pc = static_cast<float>(d1);
        qc = static_cast<float>(d0);
      }
    }
    // ---- 3: multiplicative step on the own rows (kept in av), partial norm and sum
    float s2 = 0.f, s3 = 0.f;
    if (active) {
#pragma unroll
      for (int q = 0; q < kWsCache; ++q) {
        const int i4 = tid + q * kWsThreads;
        if (4 * i4 < a.mb) {
          const float4 w = Wcol[i4];
          const float4 b = Bcol != nullptr ? Bcol[i4] : bv[q];
          const float wi[4] = {w.x, w.y, w.z, w.w};
          const float ai[4] = {av[q].x, av[q].y, av[q].z, av[q].w};
          const float bi[4] = {b.x, b.y, b.z, b.w};
          float wn[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float neg = ai[e] + wi[e] * pc;
            float pos = (kl ? bterm : bi[e]) + wi[e] * qc;
            if (powered) {
              neg = powf(neg, a.expo);
              pos = powf(pos, a.expo);
            }
            wn[e] = wi[e] * (neg / fmaxf(pos + lambda, NMFB_EPS));  // nmf.m:168
            s2 = fmaf(wn[e], wn[e], s2);
            s3 += wn[e];
          }
          av[q] = make_float4(wn[0], wn[1], wn[2], wn[3]);
        }
      }
    }
    acc[0] = s2;
    acc[1] = s3;
    block_sum<2>(acc, sh);
    NMFB_WS_STAMP(3);
    ws_exchange<N>(a.t, a.x_off + static_cast<size_t>(kMaxBlocks) * kMaxRanks * 32, static_cast<uint32_t>(a.epoch0 + round),
                   acc[0], acc[1], xs, d0, d1);
    NMFB_WS_STAMP(4);
    // ---- 4: unit L2 columns (nmf.m:169); master rows stay here, the tf32 rows go to every rank
    if (active) {
      const float mul = static_cast<float>(1.0 / sqrt(d0));
      if (tid == 0) a.wsum[k] = static_cast<double>(mul) * d1;
      float4* Wout = reinterpret_cast<float4*>(a.Wm) + col4;
      if (BULK) __syncthreads();  // everybody has taken its partial rows out of the staging buffer
#pragma unroll
      for (int q = 0; q < kWsCache; ++q) {
        const int i4 = tid + q * kWsThreads;
        if (4 * i4 < a.mb) {
          const float4 x = make_float4(av[q].x * mul, av[q].y * mul, av[q].z * mul, av[q].w * mul);
          Wout[i4] = x;
          const float4 xt = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
          if (BULK) {
            reinterpret_cast<float4*>(stage)[i4] = xt;
          } else {
#pragma unroll
            for (int r = 0; r < N; ++r) reinterpret_cast<float4*>(a.t.base[r] + a.wt_off)[col4 + i4] = xt;
          }
        }
      }
      if (BULK) {
        fence_proxy_async();  // the staged rows (generic-proxy writes) are about to be read by the TMA engine
        __syncthreads();
        if (tid == 0 && a.mb > 0) {
#pragma unroll
          for (int r = 0; r < N; ++r)
            bulk_store_1d(reinterpret_cast<float4*>(a.t.base[r] + a.wt_off) + col4, smem_u32(stage),
                          static_cast<uint32_t>(a.mb) * 4u);
          bulk_commit_group();
          bulk_wait_group0();  // the rows are written (not merely read out of shared memory) before the flag goes up
        }
        __syncthreads();  // the staging buffer is free for the next round
      }
    }
  }
  // closing: the rows every other rank owes us have landed (and ours have been delivered)
  NMFB_WS_STAMP(5);
  p2p_block_barrier(a.t, 4 * kFlagBytes, a.epoch0);
  NMFB_WS_STAMP(6);
}

// resident blocks of the kernel on one GPU (blocks spin on their peers, so the grid must fit)
template <int N, bool BULK>
inline int w_shard_capacity_t(int num_sms, size_t smem) {
  int per_sm = 1;
  if (BULK) cudaFuncSetAttribute(w_step_sharded_kernel<N, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, w_step_sharded_kernel<N, BULK>, kWsThreads, smem) != cudaSuccess ||
      per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return std::min(kMaxBlocks, per_sm * num_sms);
}
#define NMFB_WS_DISPATCH(N_, CALL)      \
  switch (N_) {                         \
    case 2: CALL(2); break;             \
    case 3: CALL(3); break;             \
    case 4: CALL(4); break;             \
    case 5: CALL(5); break;             \
    case 6: CALL(6); break;             \
    case 7: CALL(7); break;             \
    default: CALL(8); break;            \
  }
inline int w_shard_capacity(int nranks, int num_sms, bool bulk, size_t smem) {
  int cap = 1;
#define NMFB_WS_CAP(N_) cap = bulk ? w_shard_capacity_t<N_, true>(num_sms, smem) : w_shard_capacity_t<N_, false>(num_sms, 0)
  NMFB_WS_DISPATCH(nranks, NMFB_WS_CAP)
#undef NMFB_WS_CAP
  return cap;
}
inline void launch_w_step_sharded(const WShardArgs& a, int grid, bool bulk, size_t smem, cudaStream_t stream) {
#define NMFB_WS_LAUNCH(N_)                                                        \
  if (bulk) w_step_sharded_kernel<N_, true><<<grid, kWsThreads, smem, stream>>>(a); \
  else w_step_sharded_kernel<N_, false><<<grid, kWsThreads, 0, stream>>>(a)
  NMFB_WS_DISPATCH(a.t.nranks, NMFB_WS_LAUNCH)
#undef NMFB_WS_LAUNCH
}

// End of a run: every rank sends its rows of the fp32 master to all ranks (W is returned replicated).
struct WGatherArgs {
  PeerTable t;
  int K, r0, mb;
  long long ld;
  size_t wm_off;  // byte offset of the fp32 master in every rank's region
  int epoch;
};
__global__ void __launch_bounds__(kWsThreads) w_gather_rows_kernel(WGatherArgs a) {
  const float4* src = reinterpret_cast<const float4*>(a.t.base[a.t.rank] + a.wm_off);
  for (int k = blockIdx.x; k < a.K; k += gridDim.x) {
    const long long col4 = (static_cast<long long>(k) * a.ld + a.r0) >> 2;
    for (int i4 = threadIdx.x; 4 * i4 < a.mb; i4 += blockDim.x) {
      const float4 x = src[col4 + i4];
      for (int r = 0; r < a.t.nranks; ++r)
        if (r != a.t.rank) reinterpret_cast<float4*>(a.t.base[r] + a.wm_off)[col4 + i4] = x;
    }
  }
  p2p_block_barrier(a.t, 5 * kFlagBytes, a.epoch);
}

}  // namespace nmfb
