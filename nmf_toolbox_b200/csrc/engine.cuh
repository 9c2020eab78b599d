// Host-side engine shared by the algorithm drivers: the handle, device memory,
// planned GEMM operations and the small-kernel launch helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/nmfb200.h"
#include "gemm_host.cuh"

namespace nmfb {

struct Comm;  // nccl communicator wrapper (comm.cu)

inline int round_up(int x, int q) { return (x + q - 1) / q * q; }
inline long long round_up_ll(long long x, long long q) { return (x + q - 1) / q * q; }

}  // namespace nmfb

struct NmfSession;  // nmf_driver.cu

// Device blocks handed back by finished calls, kept for the next call of the same shape: a
// factorisation allocates ~4 GiB in a few dozen blocks and cudaMalloc / cudaFree of GiB-sized blocks
// cost 30-150 ms per call (measured), as much as 100 iterations at the north-star size.  Exact-size
// reuse only; bounded by `cap` (a third of the device memory; NMFB_NO_POOL=1 disables), trimmed when
// an allocation fails, and released by nmfb_trim / nmfb_destroy.
struct BlockPool {
  struct Blk {
    void* p;
    size_t bytes;
  };
  std::vector<Blk> blocks;
  size_t held = 0, cap = 0;
  void* take(size_t bytes) {
    for (size_t i = blocks.size(); i-- > 0;)
      if (blocks[i].bytes == bytes) {
        void* p = blocks[i].p;
        blocks.erase(blocks.begin() + static_cast<long>(i));
        held -= bytes;
        return p;
      }
    return nullptr;
  }
  void give(void* p, size_t bytes) {
    if (bytes > cap) {
      cudaFree(p);
      return;
    }
    blocks.push_back({p, bytes});
    held += bytes;
    while (held > cap && !blocks.empty()) {  // oldest first
      held -= blocks.front().bytes;
      cudaFree(blocks.front().p);
      blocks.erase(blocks.begin());
    }
  }
  void trim() {
    for (Blk& b : blocks) cudaFree(b.p);
    blocks.clear();
    held = 0;
  }
};

struct nmfb_handle {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: work that only depends on H runs beside the A GEMM
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  long long launches = 0;
  long long mallocs = 0;  // cudaMalloc calls that missed the block pool (nmfb_malloc_count)

  // V as given (uploaded copy or adopted device pointer)
  const float* Vraw = nullptr;
  float* Vown = nullptr;
  size_t Vown_bytes = 0;
  BlockPool pool;
  // pinned host scratch, allocated once per handle (cudaMallocHost / cudaFreeHost synchronise the
  // device and took up to 0.8 s when issued per call next to a large pinned user buffer):
  // (64 ints) [0..1] session stop flag / cost count, [2..3] run_chunked flags, [8..15] line-search state (nmfsc)
  int* pinned = nullptr;
  // two pinned bounce buffers for results that go to pageable host memory (numpy / MATLAB arrays): a direct
  // cudaMemcpy into pageable memory ran at ~3.7 GB/s (W and H of the north-star call: 9 ms of a 42 ms call)
  char* stage = nullptr;
  size_t stage_half = 0;
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  int m = 0, n = 0;
  long long ldv = 0;
  // working copy (tf32-rounded / rescaled), allocated on demand, same shape as Vraw
  float* Vwork = nullptr;
  size_t Vwork_bytes = 0;

  nmfb::Comm* comm = nullptr;
  NmfSession* sess = nullptr;
  // label constraint of the next nmf session (nmfb_constrainednmf): host pointers, consumed by the setup
  const int* tie_col2z = nullptr;
  int tie_nz = 0;
  const float* tie_Zinit = nullptr;

  // device time of the iteration loop of the last one-call entry point (nmfb_last_loop)
  double loop_ms = 0.0;
  int loop_iters = 0;
  std::vector<int> halvings;  // nmfsc: rejected line-search trials (H, W) per iteration of the last call
  // optional per-kernel timing of the two large contractions (bench.py roofline)
  bool profile = false;
  // [0] W-step GEMM, [1] H-step GEMM, [2] gram(H)+cost, [3] element-wise W step, [4] gram(W)
  std::vector<cudaEvent_t> prof_ev[5];

  int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

namespace nmfb {

#define NMFB_CUDA(h, call)                                                                    \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return (h)->fail(NMFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                   \
  } while (0)

#define NMFB_TRY(expr)          \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != NMFB_OK) return rc__; \
  } while (0)

// cudaMalloc through the handle's block pool (a failed allocation empties the pool and retries)
inline cudaError_t dev_alloc(nmfb_handle* h, void** p, size_t bytes) {
  *p = h->pool.take(bytes);
  if (*p) return cudaSuccess;
  ++h->mallocs;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess && h->pool.held > 0) {
    cudaGetLastError();
    h->pool.trim();
    e = cudaMalloc(p, bytes);
  }
  return e;
}
inline void dev_free(nmfb_handle* h, void* p, size_t bytes) {
  if (p) h->pool.give(p, bytes);
}

// Device allocations that live as long as one algorithm call / session (zero-filled).
struct Arena {
  struct Blk {
    void* p;
    size_t bytes;
  };
  std::vector<Blk> blocks;
  nmfb_handle* owner = nullptr;
  ~Arena() { release(); }
  void release() {
    if (owner && !blocks.empty()) {  // nothing of this call may still be running when the blocks are reused
      cudaStreamSynchronize(owner->stream);
      if (owner->stream2) cudaStreamSynchronize(owner->stream2);
    }
    for (Blk& b : blocks) {
      if (owner) dev_free(owner, b.p, b.bytes);
      else cudaFree(b.p);
    }
    blocks.clear();
  }
  template <class T>
  int alloc(nmfb_handle* h, T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
    if (bytes == 0) bytes = 256;
    owner = h;
    NMFB_CUDA(h, dev_alloc(h, &p, bytes));
    blocks.push_back({p, bytes});
    NMFB_CUDA(h, cudaMemsetAsync(p, 0, bytes, h->stream));
    *out = static_cast<T*>(p);
    return NMFB_OK;
  }
};

// One planned panel_gemm (tensor maps are built once; buffers never move).
struct GemmOp {
  GemmLaunch L;
  int epi = EPI_STORE;
  int splits = 1;
  float* parts = nullptr;   // split-K slabs (EPI_STORE, splits > 1)
  float* final0 = nullptr;  // where the reduced phase-0 result goes
  long long count = 0;      // elements per slab
  bool planned = false;
  mutable unsigned int sk_epoch = 0;  // launches so far of a plan with tail helpers (GemmArgs::sk_epoch)
};

struct MatRef {
  const float* base;
  long long inner, outer, pitch;
  bool mn;
};
inline GemmOperand operand(const MatRef& r) {
  return GemmOperand{{r.base, r.inner, r.outer, r.pitch}, r.mn};
}

// out0 = X0 * Y0' (+ out1 = X1 * Y1').  With allow_split the contraction of
// phase 0 may be cut over several CTAs (EPI_STORE only); the slabs are summed
// into out0 by run_gemm.
// Extra phase-0 operand pairs: acc0 = X0*Y0' + Xs[0]*Ys[0]' (+ Xs[1]*Ys[1]'), all over kdim0.
struct ExtraSegs {
  int n = 0;
  MatRef X[2];
  MatRef Y[2];
};
int plan_store(nmfb_handle* h, Arena* ar, GemmOp* op, const MatRef& X0, const MatRef& Y0,
               long long kdim0, const MatRef* X1, const MatRef* Y1, long long kdim1, int rows,
               int ncols, float* out0, float* out1, long long ldo, bool allow_split,
               const int* stop, const ExtraSegs* segs = nullptr, int tile_n = 0);
int plan_fused(nmfb_handle* h, GemmOp* op, int epi, const MatRef& X0, const MatRef& Y0,
               long long kdim0, const MatRef* X1, const MatRef* Y1, long long kdim1, int rows,
               int ncols, int ncols_valid, const int* stop, const ExtraSegs* segs = nullptr, int tile_n = 0);
int run_gemm(nmfb_handle* h, const GemmOp& op);
// Lets a planned CTA-pair launch that leaves SMs idle use them for the tails of its contractions (GemmArgs::sk_*),
// keeping `reserve_sms` SMs free for kernels meant to run beside it.  No effect when the plan is not eligible.
int enable_tail_helpers(nmfb_handle* h, Arena* ar, GemmOp* op, int reserve_sms);

// Gram matrix G = M M' of a factor stored as nvec contiguous vectors of length len
// (nvec multiple of 32): fp32 result + tf32-rounded copy.
struct GramOp {
  GemmOp g;
  float* g32 = nullptr;
  float* gtf = nullptr;  // tf32 head of g32
  float* glo = nullptr;  // tf32 tail (only when planned with a split factor)
  int nvec = 0;
};
// Mlo != nullptr: the factor is given as a tf32 head/tail pair and the Gram matrix is formed at
// fp32 accuracy from the three leading terms (hi*hi' + hi*lo' + lo*hi').
int plan_gram(nmfb_handle* h, Arena* ar, GramOp* op, const float* Mt, int nvec, int len,
              long long ld, const int* stop, const float* Mlo = nullptr, int chunk_kb = 0,
              int max_ctas = 0 /* 0 = whole GPU; else cap the grid (Gram running beside a large GEMM) */);
int run_gram(nmfb_handle* h, const GramOp& op, const int* stop, unsigned int* ticket = nullptr,
             unsigned int* gate = nullptr, unsigned int gate_value = 0);

dim3 vec_grid(int len, int nvec, int threads = 256);

// V preparation
struct VStats {
  double sumsq = 0, sum = 0, sumlog = 0;
  float vmax = 0;
  bool any_negative = false;
};
int compute_v_stats(nmfb_handle* h, bool want_log, VStats* out, double** dev_stats_out,
                    unsigned int** dev_max_out, Arena* ar);
// Vwork = tf32(V / max) ; returns sum of squares of the working copy through *sumsq_dev (device).
int prepare_v_work(nmfb_handle* h, bool divide_by_max, bool round, double* sumsq_dev,
                   const unsigned int* maxbits_dev);

int upload_colmajor(nmfb_handle* h, const float* host, int rows, int cols, float* dev, long long ld);
int download_colmajor(nmfb_handle* h, const float* dev, long long ld, int rows, int cols, float* host);
// H: host K x n column-major  <->  device row-major [Kp][ldh]
int upload_H(nmfb_handle* h, Arena* ar, const float* host, int K, int n, float* Hm, long long ldh);
int download_H(nmfb_handle* h, const float* Hm, long long ldh, int K, int n, float* host);
void fill_uniform(std::vector<float>& v, unsigned long long seed, bool clamp_eps);

int check_launch(nmfb_handle* h, const char* what);

struct WStepArgs;
// Cooperative launch of the fused W step (ew_kernels.cuh: w_step_kernel).
int launch_w_step(nmfb_handle* h, const WStepArgs& a);
struct CostArgs;
// gram GEMM + (reduce, <G_W,G_H>, cost, stop test) in one follow-up kernel.
int run_gram_cost(nmfb_handle* h, const GramOp& op, unsigned int* ticket, const CostArgs& c, bool with_cost,
                  unsigned int* gate = nullptr, unsigned int gate_value = 0);

// After an all-reduce of a Gram matrix (multi-GPU): tf32 copy, <G_W,G_H>, cost and stop test in one kernel.
int run_gram_post_allreduce(nmfb_handle* h, const GramOp& op, unsigned int* ticket, const CostArgs& c,
                            bool with_cost);

// CUDA-event pair around a group of launches when profiling is on (nmfb_profile_enable); slot meanings
// per algorithm are listed at nmfb_profile_get_all in include/nmfb200.h
inline int prof_mark(nmfb_handle* h, int which) {
  if (!h->profile) return NMFB_OK;
  cudaEvent_t e;
  NMFB_CUDA(h, cudaEventCreate(&e));
  h->prof_ev[which].push_back(e);
  NMFB_CUDA(h, cudaEventRecord(e, h->stream));
  return NMFB_OK;
}
// Brackets of the iteration loop of a one-call entry point (device time for nmfb_last_loop)
inline void loop_begin(nmfb_handle* h) {
  h->loop_ms = 0.0;
  h->loop_iters = 0;
  cudaEventRecord(h->ev0, h->stream);
}
inline void loop_end(nmfb_handle* h, int iters) {  // the stream must be idle (synchronised) when this returns
  cudaEventRecord(h->ev1, h->stream);
  cudaEventSynchronize(h->ev1);
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->loop_ms = ms;
  h->loop_iters = iters;
}

// Queue `maxiter` iterations in chunks.  The stop flag written by the cost
// kernel turns everything queued behind a converged iteration into no-ops, so
// the host only looks at the flag between chunks and never waits for the chunk
// it has just queued.  enqueue(i) queues iteration i (0-based).
template <class Fn>
int run_chunked(nmfb_handle* h, int maxiter, const int* stop_dev, Fn enqueue, int chunk = 16) {
  cudaEvent_t evs[2] = {nullptr, nullptr};
  int* flags = nullptr;
  int rc = NMFB_OK;
  for (int b = 0; b < 2 && rc == NMFB_OK; ++b)
    if (cudaEventCreateWithFlags(&evs[b], cudaEventDisableTiming) != cudaSuccess)
      rc = h->fail(NMFB_ERR_CUDA, "cudaEventCreate failed");
  flags = h->pinned + 2;
  if (rc == NMFB_OK) {
    flags[0] = flags[1] = 0;
    int c = 0, i = 0;
    while (i < maxiter && rc == NMFB_OK) {
      const int end = i + chunk < maxiter ? i + chunk : maxiter;
      for (; i < end && rc == NMFB_OK; ++i) rc = enqueue(i);
      if (rc != NMFB_OK) break;
      if (c > 0) {
        cudaEventSynchronize(evs[(c - 1) & 1]);
        if (flags[(c - 1) & 1] != 0) break;
      }
      cudaMemcpyAsync(&flags[c & 1], stop_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
      cudaEventRecord(evs[c & 1], h->stream);
      ++c;
    }
  }
  for (int b = 0; b < 2; ++b)
    if (evs[b]) cudaEventDestroy(evs[b]);
  cudaStreamSynchronize(h->stream);  // the last flag copy must not land after the scratch is reused
  return rc;
}

}  // namespace nmfb
