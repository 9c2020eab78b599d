// Engine helpers: planned GEMMs, Gram matrices, V preparation, host<->device
// transfers in the reference's (MATLAB, column-major) layouts.
#include "engine.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ew_kernels.cuh"

namespace nmfb {

int check_launch(nmfb_handle* h, const char* what) {
  ++h->launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return h->fail(NMFB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return NMFB_OK;
}

dim3 vec_grid(int len, int nvec, int threads) {
  // 4 elements per thread: these kernels are latency-bound streams over K x n / m x K arrays, and
  // with few vectors (cnmf: K = 64) a coarser grid left most of the GPU idle
  int bx = (len + threads * 4 - 1) / (threads * 4);
  bx = std::max(1, std::min(bx, 64));
  return dim3(bx, nvec, 1);
}

static int plan_common(nmfb_handle* h, GemmOp* op, const MatRef& X0, const MatRef& Y0,
                       long long kdim0, const MatRef* X1, const MatRef* Y1, long long kdim1,
                       int rows, int ncols, int splits_hint, const ExtraSegs* segs = nullptr, int tile_n = 0) {
  GemmOperand x0 = operand(X0), y0 = operand(Y0), x1{}, y1{};
  if (X1) {
    x1 = operand(*X1);
    y1 = operand(*Y1);
  }
  const int cg = choose_cg(op->epi, rows, ncols, Y0.mn, Y1 ? Y1->mn : false, tile_n);
  std::string e = plan_gemm(&op->L, x0, y0, kdim0, X1 ? &x1 : nullptr, X1 ? &y1 : nullptr, kdim1,
                            rows, ncols, splits_hint, h->num_sms, cg, tile_n);
  if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "plan_gemm: %s", e.c_str());
  for (int sgi = 0; segs && sgi < segs->n; ++sgi) {
    e = add_segment(&op->L, sgi + 1, operand(segs->X[sgi]), operand(segs->Y[sgi]), h->num_sms, splits_hint);
    if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "add_segment: %s", e.c_str());
  }
  op->splits = static_cast<int>(op->L.grid.z);
  op->planned = true;
  return NMFB_OK;
}

int plan_store(nmfb_handle* h, Arena* ar, GemmOp* op, const MatRef& X0, const MatRef& Y0,
               long long kdim0, const MatRef* X1, const MatRef* Y1, long long kdim1, int rows,
               int ncols, float* out0, float* out1, long long ldo, bool allow_split,
               const int* stop, const ExtraSegs* segs, int tile_n) {
  op->epi = EPI_STORE;
  NMFB_TRY(plan_common(h, op, X0, Y0, kdim0, X1, Y1, kdim1, rows, ncols, allow_split ? 0 : 1, segs, tile_n));
  GemmArgs& a = op->L.args;
  a.stop = stop;
  a.out1 = out1;
  a.ldo = ldo;
  op->final0 = out0;
  op->count = static_cast<long long>(ncols) * ldo;
  if (op->splits > 1) {
    NMFB_TRY(ar->alloc(h, &op->parts, static_cast<size_t>(op->splits) * op->count));
    a.out0 = op->parts;
    a.split_stride = op->count;
  } else {
    a.out0 = out0;
    a.split_stride = 0;
  }
  return NMFB_OK;
}

int plan_fused(nmfb_handle* h, GemmOp* op, int epi, const MatRef& X0, const MatRef& Y0,
               long long kdim0, const MatRef* X1, const MatRef* Y1, long long kdim1, int rows,
               int ncols, int ncols_valid, const int* stop, const ExtraSegs* segs, int tile_n) {
  op->epi = epi;
  NMFB_TRY(plan_common(h, op, X0, Y0, kdim0, X1, Y1, kdim1, rows, ncols, 1, segs, tile_n));
  op->L.args.stop = stop;
  op->L.args.ncols_valid = ncols_valid;
  return NMFB_OK;
}

int enable_tail_helpers(nmfb_handle* h, Arena* ar, GemmOp* op, int reserve_sms) {
  if (!op->planned || op->splits != 1) return NMFB_OK;
  int kp = 0;
  const int helpers = plan_tail_helpers(op->L, op->epi, h->num_sms, reserve_sms, &kp);
  if (helpers <= 0) return NMFB_OK;
  const size_t tiles = op->L.grid.x / 2;
  float* part = nullptr;
  unsigned int* flags = nullptr;
  NMFB_TRY(ar->alloc(h, &part, tiles * static_cast<size_t>(op->L.args.ncols) * (2 * kTileM)));
  NMFB_TRY(ar->alloc(h, &flags, 2 * tiles));
  set_tail_helpers(&op->L, helpers, kp, part, flags);
  op->sk_epoch = 0;
  return NMFB_OK;
}

int run_gemm(nmfb_handle* h, const GemmOp& op) {
  if (!op.planned) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "internal: GEMM not planned");
  if (op.L.args.sk_helpers > 0) const_cast<GemmArgs&>(op.L.args).sk_epoch = ++op.sk_epoch;  // flags count launches
  std::string e = launch_gemm(op.L, op.epi, h->stream);
  ++h->launches;
  if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "%s", e.c_str());
  if (op.epi == EPI_STORE && op.splits > 1) {
    const int threads = 256;
    const int blocks = static_cast<int>(std::min<long long>((op.count + threads - 1) / threads, 4096));
    split_reduce_kernel<<<blocks, threads, 0, h->stream>>>(op.parts, op.splits, op.count, op.final0,
                                                           op.count, op.L.args.stop);
    NMFB_TRY(check_launch(h, "split_reduce"));
  }
  return NMFB_OK;
}

int plan_gram(nmfb_handle* h, Arena* ar, GramOp* op, const float* Mt, int nvec, int len,
              long long ld, const int* stop, const float* Mlo, int chunk_kb, int max_ctas) {
  op->nvec = nvec;
  NMFB_TRY(ar->alloc(h, &op->g32, static_cast<size_t>(nvec) * nvec));
  NMFB_TRY(ar->alloc(h, &op->gtf, static_cast<size_t>(nvec) * nvec));
  MatRef M{Mt, len, nvec, ld, false};
  op->g.epi = EPI_STORE;
  ExtraSegs segs;
  if (Mlo) {
    NMFB_TRY(ar->alloc(h, &op->glo, static_cast<size_t>(nvec) * nvec));
    MatRef L{Mlo, len, nvec, ld, false};
    segs.n = 2;
    segs.X[0] = M;
    segs.Y[0] = L;
    segs.X[1] = L;
    segs.Y[1] = M;
  }
  int splits_hint = 0;
  if (max_ctas > 0) {
    const int cg = pair_eligible(nvec, nvec, false, false) ? 2 : 1;
    const int tiles = (nvec + kTileM * cg - 1) / (kTileM * cg) * cg * ((nvec + kMaxN - 1) / kMaxN);
    splits_hint = std::max(2, max_ctas / std::max(1, tiles));
  }
  NMFB_TRY(plan_common(h, &op->g, M, M, len, nullptr, nullptr, 0, nvec, nvec, splits_hint, Mlo ? &segs : nullptr));
  op->g.L.args.chunk_kb = chunk_kb;
  GemmArgs& a = op->g.L.args;
  a.stop = stop;
  a.ldo = nvec;
  op->g.count = static_cast<long long>(nvec) * nvec;
  NMFB_TRY(ar->alloc(h, &op->g.parts, static_cast<size_t>(op->g.splits) * op->g.count));
  a.out0 = op->g.parts;
  a.split_stride = op->g.count;
  return NMFB_OK;
}

int run_gram(nmfb_handle* h, const GramOp& op, const int* stop, unsigned int* ticket, unsigned int* gate,
             unsigned int gate_value) {
  std::string e = launch_gemm(op.g.L, EPI_STORE, h->stream);
  ++h->launches;
  if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "%s", e.c_str());
  const int count = op.nvec * op.nvec;
  gram_reduce_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(op.g.parts, op.g.splits, count,
                                                                 op.g32, op.gtf, op.glo, count, stop, ticket, gate,
                                                                 gate_value);
  return check_launch(h, "gram_reduce");
}

int run_gram_cost(nmfb_handle* h, const GramOp& op, unsigned int* ticket, const CostArgs& c, bool with_cost,
                  unsigned int* gate, unsigned int gate_value) {
  std::string e = launch_gemm(op.g.L, EPI_STORE, h->stream);
  ++h->launches;
  if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "%s", e.c_str());
  const int count = op.nvec * op.nvec;
  gram_reduce_cost_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(op.g.parts, op.g.splits, count, op.g32,
                                                                      op.gtf, count, ticket, c, with_cost ? 1 : 0,
                                                                      gate, gate_value);
  return check_launch(h, "gram_reduce_cost");
}

int run_gram_post_allreduce(nmfb_handle* h, const GramOp& op, unsigned int* ticket, const CostArgs& c,
                            bool with_cost) {
  const int count = op.nvec * op.nvec;  // "one slab" = the reduced matrix itself
  gram_reduce_cost_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(op.g32, 1, count, op.g32, op.gtf, count,
                                                                      ticket, c, with_cost ? 1 : 0, nullptr, 0);
  return check_launch(h, "gram_post_allreduce");
}

int launch_w_step(nmfb_handle* h, const WStepArgs& a) {
  if (a.m <= 256 * kWCache)
    w_step_kernel<true, 256><<<a.K, 256, 0, h->stream>>>(a);
  else if (a.m <= kWThreads * kWCache)
    w_step_kernel<true, kWThreads><<<a.K, kWThreads, 0, h->stream>>>(a);
  else
    w_step_kernel<false, kWThreads><<<a.K, kWThreads, 0, h->stream>>>(a);
  return check_launch(h, "w_step");
}

// ---------------------------------------------------------------- V
int compute_v_stats(nmfb_handle* h, bool want_log, VStats* out, double** dev_stats_out,
                    unsigned int** dev_max_out, Arena* ar) {
  double* st = nullptr;
  unsigned int* su = nullptr;
  NMFB_TRY(ar->alloc(h, &st, 4));
  NMFB_TRY(ar->alloc(h, &su, 2));
  const int blocks = std::min(h->n, h->num_sms * 8);
  v_stats_kernel<<<blocks, 256, 0, h->stream>>>(h->Vraw, h->m, h->n, h->ldv, st, su, want_log ? 1 : 0);
  NMFB_TRY(check_launch(h, "v_stats"));
  double hs[4];
  unsigned int hu[2];
  NMFB_CUDA(h, cudaMemcpyAsync(hs, st, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaMemcpyAsync(hu, su, sizeof(hu), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  out->sumsq = hs[0];
  out->sum = hs[1];
  out->sumlog = hs[2];
  std::memcpy(&out->vmax, &hu[0], 4);
  out->any_negative = hu[1] != 0;
  if (dev_stats_out) *dev_stats_out = st;
  if (dev_max_out) *dev_max_out = su;
  return NMFB_OK;
}

int prepare_v_work(nmfb_handle* h, bool divide_by_max, bool round, double* sumsq_dev,
                   const unsigned int* maxbits_dev) {
  const size_t bytes = static_cast<size_t>(h->n) * h->ldv * sizeof(float);
  if (h->Vwork == nullptr || h->Vwork_bytes != bytes) {  // through the block pool: trimmed and retried on failure
    if (h->Vwork) {
      cudaStreamSynchronize(h->stream);
      dev_free(h, h->Vwork, h->Vwork_bytes);
    }
    h->Vwork = nullptr;
    h->Vwork_bytes = 0;
    void* p = nullptr;
    NMFB_CUDA(h, dev_alloc(h, &p, bytes));
    h->Vwork = static_cast<float*>(p);
    h->Vwork_bytes = bytes;
  }
  const int blocks = std::min(h->n, h->num_sms * 8);
  v_prepare_kernel<<<blocks, 256, 0, h->stream>>>(h->Vraw, h->Vwork, h->m, h->n, h->ldv, h->ldv,
                                                  divide_by_max ? maxbits_dev : nullptr,
                                                  round ? 1 : 0, sumsq_dev);
  return check_launch(h, "v_prepare");
}

// ---------------------------------------------------------------- transfers
int upload_colmajor(nmfb_handle* h, const float* host, int rows, int cols, float* dev, long long ld) {
  NMFB_CUDA(h, cudaMemcpy2DAsync(dev, ld * sizeof(float), host, static_cast<size_t>(rows) * sizeof(float),
                                 static_cast<size_t>(rows) * sizeof(float), cols,
                                 cudaMemcpyHostToDevice, h->stream));
  return NMFB_OK;
}
// Device -> host copy of `nrows` rows of `width` bytes (device pitch `pitch`, host rows packed).  Pinned or
// registered destinations are written directly; pageable ones through the handle's two pinned bounce buffers, the
// DMA of one chunk running while the host copies the previous one out.  Returns with the data in `host`.
static int d2h_rows(nmfb_handle* h, char* host, const char* dev, size_t pitch, size_t width, size_t nrows) {
  if (nrows == 0 || width == 0) return NMFB_OK;
  cudaPointerAttributes attr{};
  const bool pageable = cudaPointerGetAttributes(&attr, host) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  const char* env = std::getenv("NMFB_NO_STAGING");
  if (!pageable || h->stage == nullptr || width == 0 || width > h->stage_half || (env && env[0] == '1')) {
    NMFB_CUDA(h, cudaMemcpy2DAsync(host, width, dev, pitch, width, nrows, cudaMemcpyDeviceToHost, h->stream));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return NMFB_OK;
  }
  const size_t rpc = std::max<size_t>(1, h->stage_half / width);  // rows per chunk
  const size_t nchunks = (nrows + rpc - 1) / rpc;
  auto issue = [&](size_t c) -> cudaError_t {
    const int b = static_cast<int>(c & 1);
    const size_t r0 = c * rpc, r = std::min(rpc, nrows - r0);
    cudaError_t e = cudaMemcpy2DAsync(h->stage + b * h->stage_half, width, dev + r0 * pitch, pitch, width, r,
                                      cudaMemcpyDeviceToHost, h->stream);
    return e != cudaSuccess ? e : cudaEventRecord(h->ev_stage[b], h->stream);
  };
  NMFB_CUDA(h, issue(0));
  if (nchunks > 1) NMFB_CUDA(h, issue(1));
  for (size_t c = 0; c < nchunks; ++c) {
    const int b = static_cast<int>(c & 1);
    NMFB_CUDA(h, cudaEventSynchronize(h->ev_stage[b]));
    std::memcpy(host + c * rpc * width, h->stage + b * h->stage_half, std::min(rpc, nrows - c * rpc) * width);
    if (c + 2 < nchunks) NMFB_CUDA(h, issue(c + 2));  // this buffer is free again
  }
  return NMFB_OK;
}

int download_colmajor(nmfb_handle* h, const float* dev, long long ld, int rows, int cols, float* host) {
  return d2h_rows(h, reinterpret_cast<char*>(host), reinterpret_cast<const char*>(dev), ld * sizeof(float),
                  static_cast<size_t>(rows) * sizeof(float), static_cast<size_t>(cols));
}

int upload_H(nmfb_handle* h, Arena* ar, const float* host, int K, int n, float* Hm, long long ldh) {
  float* tmp = nullptr;  // [n][K] as the host has it
  NMFB_TRY(ar->alloc(h, &tmp, static_cast<size_t>(n) * K));
  NMFB_CUDA(h, cudaMemcpyAsync(tmp, host, static_cast<size_t>(n) * K * sizeof(float),
                               cudaMemcpyHostToDevice, h->stream));
  dim3 grid((n + 31) / 32, (K + 31) / 32);
  transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(tmp, K, Hm, ldh, K, n);
  return check_launch(h, "transpose(H in)");
}
int download_H(nmfb_handle* h, const float* Hm, long long ldh, int K, int n, float* host) {
  float* tmp = nullptr;
  const size_t tmp_bytes = (static_cast<size_t>(n) * K * sizeof(float) + 255) / 256 * 256;
  {
    void* p = nullptr;
    NMFB_CUDA(h, dev_alloc(h, &p, tmp_bytes));
    tmp = static_cast<float*>(p);
  }
  dim3 grid((K + 31) / 32, (n + 31) / 32);
  // dst[r = j][c = k] = src[c = k][r = j]
  transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(Hm, ldh, tmp, K, n, K);
  int rc = check_launch(h, "transpose(H out)");
  if (rc == NMFB_OK) {  // [n][K] packed: rows of K floats
    const size_t row = static_cast<size_t>(K) * sizeof(float);
    const size_t group = std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(n), (size_t(1) << 20) / row));
    // copy in "rows" of `group` samples so that a chunk is a whole number of them (the tail separately)
    const size_t whole = static_cast<size_t>(n) / group;
    rc = d2h_rows(h, reinterpret_cast<char*>(host), reinterpret_cast<const char*>(tmp), group * row, group * row, whole);
    const size_t rest = static_cast<size_t>(n) - whole * group;
    if (rc == NMFB_OK && rest > 0)
      rc = d2h_rows(h, reinterpret_cast<char*>(host) + whole * group * row,
                    reinterpret_cast<const char*>(tmp) + whole * group * row, rest * row, rest * row, 1);
  }
  cudaStreamSynchronize(h->stream);  // (also on the error paths) nothing may still write the block
  dev_free(h, tmp, tmp_bytes);
  return rc;
}

// splitmix64 -> uniform (0,1); stands in for MATLAB's rand (nmf.m:277,298).
void fill_uniform(std::vector<float>& v, unsigned long long seed, bool clamp_eps) {
  unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  for (float& x : v) {
    s += 0x9E3779B97F4A7C15ull;
    unsigned long long z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    float u = static_cast<float>((z >> 40) + 1) * (1.0f / 16777218.0f);
    if (clamp_eps) u = std::max(u, 2.220446049250313e-16f);
    x = u;
  }
}

}  // namespace nmfb
