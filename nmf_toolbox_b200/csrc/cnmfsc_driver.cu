// cnmfsc driver: convolutive NMF with sparseness constraints, cnmfsc.m:66-277,
// for W_sparsity == 0 (see the note on the W line search at the end of this comment).
//
// Stacked form as in cnmf_driver.cu: Wc = [W_1 ... W_T] (m x KT), Hs = [H_1; ...; H_T] with H_t = H
// shifted right by t-1 columns, V_hat = Wc*Hs.  One iteration (cnmfsc.m:155-276):
//   H step   neg = sum_t W_t' shift<-(V, t-1)     = fold(Wc' V)
//            pos = sum_t W_t' shift<-(V_hat, t-1) = fold((Wc'Wc) Hs)          (cnmfsc.m:160-165)
//            H_sparsity > 0: projected gradient step with the reference's line search on explicitly
//            evaluated objectives 0.5*|V - Wc*Hs(Hnew)|^2 (166-199); else the multiplicative step,
//            rows of H to unit L2 and every frame of W scaled by the norms (202-209)
//   W step   frames in order, each from the V_hat that already contains the frames updated before it
//            (257-263): pos_t = V_hat Hs_t' = Wc_current * (Hs Hs')(:, frame t), W_t <- W_t .* neg_t ./
//            max(pos_t, eps).  (The clamp max(V_hat + ..., 0) of line 262 never acts on non-negative
//            factors up to rounding and is not reproduced.)
//   cost     0.5*|V - Wc*Hs|^2, stop rule of lines 273-276.
// All contractions are split-tf32 products (three operand segments, short accumulation chunks) for the
// reason given in nmfsc_driver.cu: the line search compares objectives that differ by parts in 1e5.
//
// W_sparsity > 0 (cnmfsc.m:100-110, 216-254) is reproduced literally, quirks included: the initial
// projection is applied to W while the iterations start from the unprojected W0 (W0 = W only at the end
// of an iteration, line 266), so the first cost and the first H gradient see V_hat = RFD(W, H) with
// the projected W (a cross Gram matrix W0c' Wc on the device); and the W line search reconstructs its
// trial from the single frame alone (line 235: RFD(Wnew, H) with a 2-D Wnew is Wnew*H, unshifted),
// compares that with the objective it inherited (line 218) and hands this V_hat to the next frame.
// On ordinary data the trial cannot win and the function returns by step-size underflow
// ("Algorithm converged", lines 245-249) after ~665 halvings with the cost trimmed.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "comm.cuh"
#include "engine.cuh"
#include "ew_kernels.cuh"

using namespace nmfb;

namespace cnmfscdetail {

// Hs[k + K*t][j] = H[k][j - t] for j >= t, else 0 (unrounded; split into head/tail afterwards)
__global__ void stack_kernel(const float* __restrict__ H, float* __restrict__ Hs, int K, int T, int n, long long ld) {
  const int c = blockIdx.y;
  const int k = c % K, t = c / K;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    Hs[c * ld + j] = (j >= t) ? H[k * ld + (j - t)] : 0.f;
}

// out[k][j] = sum_t P[k + K*t][j + t]   (cnmfsc.m:160-165)
__global__ void fold_kernel(const float* __restrict__ P, float* __restrict__ out, int K, int T, int n, long long ld) {
  const int k = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T && j + t < n; ++t) s += P[static_cast<long long>(k + K * t) * ld + j + t];
    out[static_cast<long long>(k) * ld + j] = s;
  }
}

constexpr int kChunk = 2;  // k-blocks per TMEM accumulation chunk, as in nmfsc

struct State {
  Arena ar;
  int K = 0, T = 0, KT = 0, KTp = 0, Kp = 0, m = 0, n = 0;
  long long ldw = 0, ldh = 0;
  float *Wm = nullptr, *Wt = nullptr, *Wl = nullptr, *Wsave = nullptr;  // W0 master, head, tail; copy for W_fixed
  float *Hm = nullptr, *Hnew = nullptr;                                  // K x n
  float *Hs = nullptr, *Hst = nullptr, *Hsl = nullptr;                   // stack of the current H (raw, head, tail)
  float *Hn = nullptr, *Hnt = nullptr, *Hnl = nullptr;                   // stack of the trial H
  float *Vhi = nullptr, *Vlo = nullptr;
  float *P = nullptr, *D = nullptr, *Nf = nullptr, *Df = nullptr, *A = nullptr, *Bt = nullptr;
  double *scal = nullptr, *sq = nullptr;
  int* fail = nullptr;
  GramOp gramW, gramH;
  GemmOp gemmN, gemmD, gemmA, residCur, residTrial;
  std::vector<GemmOp> gemmBt;
  ResidOp rsCur, rsTrial;
  bool fused_resid = false;
  // W_sparsity > 0: the returned W (projected at the start), its head/tail, the frame trial, the cross Gram matrix
  float *Wret = nullptr, *Wrt = nullptr, *Wrl = nullptr, *Wnew = nullptr, *Wnt = nullptr, *Wnl = nullptr;
  float *Cx = nullptr, *Cxt = nullptr, *Cxl = nullptr;
  GemmOp gemmC, gemmD1, residInit, residW;
  std::vector<GemmOp> gemmPosT;
  ResidOp rsInit, rsW;
  bool fused_w = false;
};

int objective(nmfb_handle* h, State* s, const GemmOp& op, const ResidOp& rs, double* out) {
  NMFB_CUDA(h, cudaMemsetAsync(s->scal, 0, 2 * sizeof(double), h->stream));
  if (s->fused_resid) NMFB_TRY(run_resid(h, rs));
  else NMFB_TRY(run_gemm(h, op));
  double v[2];
  int failed = 0;
  NMFB_CUDA(h, cudaMemcpyAsync(v, s->scal, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaMemcpyAsync(&failed, s->fail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (failed) return h->fail(NMFB_ERR_PROJFUNC, "projfunc diverged (non-finite values)");
  *out = 0.5 * v[0];  // cnmfsc.m:153,181,270
  return NMFB_OK;
}

int split_to(nmfb_handle* h, const float* src, float* hi, float* lo, int nvec, int len, long long ld) {
  dim3 grid = vec_grid(len, nvec);
  grid.y = std::min<unsigned>(grid.y, 8192u);
  split_copy_kernel<<<grid, 256, 0, h->stream>>>(src, hi, lo, nvec, len, ld);
  return check_launch(h, "split_copy");
}

// Hs (raw), head and tail from a K x n matrix
int stack_split(nmfb_handle* h, State* s, const float* H, float* raw, float* hi, float* lo) {
  stack_kernel<<<vec_grid(s->n, s->KT), 256, 0, h->stream>>>(H, raw, s->K, s->T, s->n, s->ldh);
  NMFB_TRY(check_launch(h, "stack"));
  return split_to(h, raw, hi, lo, s->KT, s->n, s->ldh);
}

int project_rows(nmfb_handle* h, State* s, float* X, double k1) {
  if (s->n > kProjThreads * 32 * kProjMaskWords)
    return h->fail(NMFB_ERR_UNSUPPORTED, "cnmfsc: projfunc vectors longer than %d are not supported",
                   kProjThreads * 32 * kProjMaskWords);
  launch_projfunc(h->stream, s->K, X, s->n, s->ldh, k1, 1.0, 1, nullptr, s->fail);
  return check_launch(h, "projfunc");
}

int run(nmfb_handle* h, State* s, int K, int T, const nmfb_config* cfg_in, float* W_out, float* H_out,
        double* cost_out, int* n_cost) {
  if (h->Vraw == nullptr) return h->fail(NMFB_ERR_NO_DATA, "cnmfsc: call nmfb_set_V first");
  if (K <= 0 || T <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "cnmfsc: K and context_len must be positive");
  if (comm_size(h->comm) > 1) return h->fail(NMFB_ERR_UNSUPPORTED, "cnmfsc: single GPU only");
  nmfb_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  if (cfg_in) cfg = *cfg_in;
  if (cfg.maxiter <= 0) cfg.maxiter = 100;         // cnmfsc.m:137-139
  if (!(cfg.tolerance > 0)) cfg.tolerance = 1e-3;  // cnmfsc.m:142-144
  double sH = cfg.H_sparsity;
  if (sH > 1) sH = 1;  // cnmfsc.m:117-119
  double sW = cfg.W_sparsity;
  if (sW > 1) sW = 1;  // cnmfsc.m:101-103
  const bool spW = sW > 0;
  const bool W_fixed = cfg.W_fixed != 0, H_fixed = cfg.H_fixed != 0;
  const int m = h->m, n = h->n;
  s->K = K;
  s->T = T;
  s->KT = K * T;
  s->KTp = round_up(s->KT, 32);
  s->Kp = round_up(K, 32);
  s->m = m;
  s->n = n;
  s->ldw = round_up(m, 4);
  s->ldh = round_up(n, 4);
  const int KT = s->KT, KTp = s->KTp, Kp = s->Kp;
  const long long ldw = s->ldw, ldh = s->ldh;
  Arena* ar = &s->ar;

  // cnmfsc.m:67-72: reject negative data, rescale by the maximum
  VStats st;
  unsigned int* maxbits = nullptr;
  NMFB_TRY(compute_v_stats(h, false, &st, nullptr, &maxbits, ar));
  if (st.any_negative) return h->fail(NMFB_ERR_NEGATIVE_DATA, "Negative values in data!");
  NMFB_TRY(prepare_v_work(h, true, false, nullptr, maxbits));  // Vwork = V / max(V), unrounded
  NMFB_TRY(ar->alloc(h, &s->Vhi, static_cast<size_t>(n) * h->ldv));
  NMFB_TRY(ar->alloc(h, &s->Vlo, static_cast<size_t>(n) * h->ldv));
  NMFB_TRY(split_to(h, h->Vwork, s->Vhi, s->Vlo, n, static_cast<int>(h->ldv), h->ldv));

  const size_t cw = static_cast<size_t>(KTp) * ldw, chs = static_cast<size_t>(KTp) * ldh;
  const size_t ch = static_cast<size_t>(Kp) * ldh;
  NMFB_TRY(ar->alloc(h, &s->Wm, cw));
  NMFB_TRY(ar->alloc(h, &s->Wt, cw));
  NMFB_TRY(ar->alloc(h, &s->Wl, cw));
  if (W_fixed) NMFB_TRY(ar->alloc(h, &s->Wsave, cw));
  if (spW) {
    NMFB_TRY(ar->alloc(h, &s->Wret, cw));
    NMFB_TRY(ar->alloc(h, &s->Wrt, cw));
    NMFB_TRY(ar->alloc(h, &s->Wrl, cw));
    NMFB_TRY(ar->alloc(h, &s->Wnew, static_cast<size_t>(Kp) * ldw));
    NMFB_TRY(ar->alloc(h, &s->Wnt, static_cast<size_t>(Kp) * ldw));
    NMFB_TRY(ar->alloc(h, &s->Wnl, static_cast<size_t>(Kp) * ldw));
    NMFB_TRY(ar->alloc(h, &s->Cx, static_cast<size_t>(KTp) * KTp));
    NMFB_TRY(ar->alloc(h, &s->Cxt, static_cast<size_t>(KTp) * KTp));
    NMFB_TRY(ar->alloc(h, &s->Cxl, static_cast<size_t>(KTp) * KTp));
  }
  NMFB_TRY(ar->alloc(h, &s->A, cw));
  NMFB_TRY(ar->alloc(h, &s->Bt, static_cast<size_t>(Kp) * ldw));
  NMFB_TRY(ar->alloc(h, &s->Hm, ch));
  NMFB_TRY(ar->alloc(h, &s->Hnew, ch));
  NMFB_TRY(ar->alloc(h, &s->Nf, ch));
  NMFB_TRY(ar->alloc(h, &s->Df, ch));
  NMFB_TRY(ar->alloc(h, &s->Hs, chs));
  NMFB_TRY(ar->alloc(h, &s->Hst, chs));
  NMFB_TRY(ar->alloc(h, &s->Hsl, chs));
  NMFB_TRY(ar->alloc(h, &s->Hn, chs));
  NMFB_TRY(ar->alloc(h, &s->Hnt, chs));
  NMFB_TRY(ar->alloc(h, &s->Hnl, chs));
  NMFB_TRY(ar->alloc(h, &s->P, chs));
  NMFB_TRY(ar->alloc(h, &s->D, chs));
  NMFB_TRY(ar->alloc(h, &s->scal, 2));
  NMFB_TRY(ar->alloc(h, &s->sq, Kp));
  NMFB_TRY(ar->alloc(h, &s->fail, 1));

  {  // cnmfsc.m:83-95
    std::vector<float> tmp;
    const float* Wsrc = cfg.W_init;
    if (!Wsrc) {
      tmp.resize(static_cast<size_t>(m) * KT);
      fill_uniform(tmp, cfg.seed * 2 + 1, false);
      Wsrc = tmp.data();
    }
    NMFB_TRY(upload_colmajor(h, Wsrc, m, KT, s->Wm, ldw));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    const float* Hsrc = cfg.H_init;
    if (!Hsrc) {
      tmp.resize(static_cast<size_t>(K) * n);
      fill_uniform(tmp, cfg.seed * 2 + 2, false);
      Hsrc = tmp.data();
    }
    NMFB_TRY(upload_H(h, ar, Hsrc, K, n, s->Hm, ldh));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!cfg.H_init) {  // default H_init has unit-L2 rows (cnmfsc.m:90)
      vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, ldh, nullptr, s->sq, nullptr);
      NMFB_TRY(check_launch(h, "vec_sums"));
      renorm_pair_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, n, ldh, s->Wm, 0, ldw, s->sq);
      NMFB_TRY(check_launch(h, "renorm"));
    }
  }
  double L1s = 0;
  if (sH > 0) {  // cnmfsc.m:116-124
    L1s = std::sqrt(static_cast<double>(n)) - (std::sqrt(static_cast<double>(n)) - 1) * sH;
    NMFB_TRY(project_rows(h, s, s->Hm, L1s));
  }
  double L1a = 0;
  if (spW) {  // cnmfsc.m:100-110: W is projected, W0 is not
    if (m > kProjThreads * 32 * kProjMaskWords)
      return h->fail(NMFB_ERR_UNSUPPORTED, "cnmfsc: projfunc vectors longer than %d are not supported",
                     kProjThreads * 32 * kProjMaskWords);
    L1a = std::sqrt(static_cast<double>(m)) - (std::sqrt(static_cast<double>(m)) - 1) * sW;
    NMFB_CUDA(h, cudaMemcpyAsync(s->Wret, s->Wm, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    launch_projfunc(h->stream, KT, s->Wret, m, ldw, L1a, 1.0, 1, nullptr, s->fail);
    NMFB_TRY(check_launch(h, "projfunc(W init)"));
    NMFB_TRY(split_to(h, s->Wret, s->Wrt, s->Wrl, KT, m, ldw));
  }
  if (W_fixed)  // the W that "W0 = W" (cnmfsc.m:266) restores at the end of every iteration
    NMFB_CUDA(h, cudaMemcpyAsync(s->Wsave, spW ? s->Wret : s->Wm, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, KT, m, ldw));
  NMFB_TRY(stack_split(h, s, s->Hm, s->Hs, s->Hst, s->Hsl));

  // ---- contractions (split-tf32: acc = Xhi*Yhi' + Xlo*Yhi' + Xhi*Ylo')
  auto three = [](const MatRef& Xhi, const MatRef& Xlo, const MatRef& Yhi, const MatRef& Ylo) {
    ExtraSegs e;
    e.n = 2;
    e.X[0] = Xlo;
    e.Y[0] = Yhi;
    e.X[1] = Xhi;
    e.Y[1] = Ylo;
    return e;
  };
  NMFB_TRY(plan_gram(h, ar, &s->gramW, s->Wt, KTp, m, ldw, nullptr, s->Wl, kChunk));
  NMFB_TRY(plan_gram(h, ar, &s->gramH, s->Hst, KTp, n, ldh, nullptr, s->Hsl, kChunk));
  {
    // P = Wc' V: rows = columns j of V (K-major), contraction over i
    MatRef Vk_hi{s->Vhi, m, n, h->ldv, false}, Vk_lo{s->Vlo, m, n, h->ldv, false};
    MatRef Wk_hi{s->Wt, m, KTp, ldw, false}, Wk_lo{s->Wl, m, KTp, ldw, false};
    ExtraSegs eN = three(Vk_hi, Vk_lo, Wk_hi, Wk_lo);
    const int tiles_h = (n + kTileM - 1) / kTileM * ((KTp + kMaxN - 1) / kMaxN);
    NMFB_TRY(plan_store(h, ar, &s->gemmN, Vk_hi, Wk_hi, m, nullptr, nullptr, 0, n, KTp, s->P, nullptr, ldh,
                        tiles_h * 2 <= h->num_sms, nullptr, &eN));
    // D = (Wc'Wc) Hs: rows j (Hs is MN-major there), contraction over the stacked index
    MatRef Hm_hi{s->Hst, n, KTp, ldh, true}, Hm_lo{s->Hsl, n, KTp, ldh, true};
    MatRef Gw_hi{s->gramW.gtf, KTp, KTp, KTp, false}, Gw_lo{s->gramW.glo, KTp, KTp, KTp, false};
    ExtraSegs eD = three(Hm_hi, Hm_lo, Gw_hi, Gw_lo);
    NMFB_TRY(plan_store(h, ar, &s->gemmD, Hm_hi, Gw_hi, KTp, nullptr, nullptr, 0, n, KTp, s->D, nullptr, ldh, false,
                        nullptr, &eD));
    // A = V Hs': rows i of V (MN-major), contraction over j
    MatRef Vm_hi{s->Vhi, m, n, h->ldv, true}, Vm_lo{s->Vlo, m, n, h->ldv, true};
    MatRef Hk_hi{s->Hst, n, KTp, ldh, false}, Hk_lo{s->Hsl, n, KTp, ldh, false};
    ExtraSegs eA = three(Vm_hi, Vm_lo, Hk_hi, Hk_lo);
    const int tiles_w = (m + kTileM - 1) / kTileM * ((KTp + kMaxN - 1) / kMaxN);
    NMFB_TRY(plan_store(h, ar, &s->gemmA, Vm_hi, Hk_hi, n, nullptr, nullptr, 0, m, KTp, s->A, nullptr, ldw,
                        tiles_w * 2 <= h->num_sms, nullptr, &eA));
    // B_t = Wc (Hs Hs')(:, frame t): G is symmetric, so its rows tK .. tK+K-1 are the K-major operand
    MatRef Wm_hi{s->Wt, m, KTp, ldw, true}, Wm_lo{s->Wl, m, KTp, ldw, true};
    s->gemmBt.resize(T);
    for (int t = 0; t < T; ++t) {
      const size_t off = static_cast<size_t>(t) * K * KTp;
      MatRef Gh_hi{s->gramH.gtf + off, KTp, K, KTp, false}, Gh_lo{s->gramH.glo + off, KTp, K, KTp, false};
      ExtraSegs eB = three(Wm_hi, Wm_lo, Gh_hi, Gh_lo);
      NMFB_TRY(plan_store(h, ar, &s->gemmBt[t], Wm_hi, Gh_hi, KTp, nullptr, nullptr, 0, m, Kp, s->Bt, nullptr, ldw,
                          false, nullptr, &eB));
      s->gemmBt[t].L.args.chunk_kb = kChunk;
    }
    for (GemmOp* op : {&s->gemmN, &s->gemmD, &s->gemmA}) op->L.args.chunk_kb = kChunk;
    // objective 0.5*|V - Wc*Hs|^2 for the current and the trial stack
    auto plan_obj = [&](GemmOp* op, const float* Whi, const float* Wlo, int Kdim, const float* Hhi, const float* Hlo) {
      MatRef Xh{Whi, m, Kdim, ldw, true}, Xl{Wlo, m, Kdim, ldw, true};
      MatRef Yh{Hhi, n, Kdim, ldh, true}, Yl{Hlo, n, Kdim, ldh, true};
      ExtraSegs e = three(Xh, Xl, Yh, Yl);
      NMFB_TRY(plan_fused(h, op, EPI_RESID, Xh, Yh, Kdim, nullptr, nullptr, 0, m, round_up(n, 64), n, nullptr, &e));
      op->L.args.Vsrc = h->Vwork;
      op->L.args.ldv = h->ldv;
      op->L.args.scal = s->scal;
      op->L.args.chunk_kb = kChunk;
      std::string pe = set_v_prefetch(&op->L, h->Vwork, m, n, h->ldv);
      if (!pe.empty()) return h->fail(NMFB_ERR_CUDA, "%s", pe.c_str());
      return static_cast<int>(NMFB_OK);
    };
    s->fused_resid = KTp <= kKlMaxKp && std::getenv("NMFB_RESID_UNFUSED") == nullptr;
    if (s->fused_resid) {
      NMFB_TRY(nmfb::plan_resid(h, &s->rsCur, s->Wt, s->Wl, ldw, s->Hst, s->Hsl, ldh, h->Vwork, h->ldv, m, n, KTp, s->scal));
      NMFB_TRY(nmfb::plan_resid(h, &s->rsTrial, s->Wt, s->Wl, ldw, s->Hnt, s->Hnl, ldh, h->Vwork, h->ldv, m, n, KTp, s->scal));
    } else {
      NMFB_TRY(plan_obj(&s->residCur, s->Wt, s->Wl, KTp, s->Hst, s->Hsl));
      NMFB_TRY(plan_obj(&s->residTrial, s->Wt, s->Wl, KTp, s->Hnt, s->Hnl));
    }
    if (spW) {
      // initial cost with the projected W (cnmfsc.m:152-153)
      if (s->fused_resid) NMFB_TRY(nmfb::plan_resid(h, &s->rsInit, s->Wrt, s->Wrl, ldw, s->Hst, s->Hsl, ldh, h->Vwork, h->ldv, m, n, KTp, s->scal));
      else NMFB_TRY(plan_obj(&s->residInit, s->Wrt, s->Wrl, KTp, s->Hst, s->Hsl));
      // frame trial: 0.5*|V - Wnew*H|^2 with the unshifted H = the first K rows of the stack (line 235);
      // rows K..Kp of the stack meet the zero padding columns of Wnew
      s->fused_w = Kp <= kKlMaxKp && std::getenv("NMFB_RESID_UNFUSED") == nullptr;
      if (s->fused_w) NMFB_TRY(nmfb::plan_resid(h, &s->rsW, s->Wnt, s->Wnl, ldw, s->Hst, s->Hsl, ldh, h->Vwork, h->ldv, m, n, Kp, s->scal));
      else NMFB_TRY(plan_obj(&s->residW, s->Wnt, s->Wnl, Kp, s->Hst, s->Hsl));
      // cross Gram C[c][c'] = sum_i W0[i][c] W[i][c'] for the first H gradient: W0c' V_hat = C Hs
      MatRef Xr_hi{s->Wrt, m, KTp, ldw, false}, Xr_lo{s->Wrl, m, KTp, ldw, false};
      MatRef Y0_hi{s->Wt, m, KTp, ldw, false}, Y0_lo{s->Wl, m, KTp, ldw, false};
      ExtraSegs eC = three(Xr_hi, Xr_lo, Y0_hi, Y0_lo);
      NMFB_TRY(plan_store(h, ar, &s->gemmC, Xr_hi, Y0_hi, m, nullptr, nullptr, 0, KTp, KTp, s->Cx, nullptr, KTp, true,
                          nullptr, &eC));
      MatRef Hm_hi{s->Hst, n, KTp, ldh, true}, Hm_lo{s->Hsl, n, KTp, ldh, true};
      MatRef C_hi{s->Cxt, KTp, KTp, KTp, false}, C_lo{s->Cxl, KTp, KTp, KTp, false};
      ExtraSegs eD1 = three(Hm_hi, Hm_lo, C_hi, C_lo);
      NMFB_TRY(plan_store(h, ar, &s->gemmD1, Hm_hi, C_hi, KTp, nullptr, nullptr, 0, n, KTp, s->D, nullptr, ldh, false,
                          nullptr, &eD1));
      // pos_t for t >= 1: (Wnew_{t-1} H) Hs_t' = Wnew_{t-1} * G[0:K, frame t]  (G symmetric: rows of frame t, first K entries)
      MatRef Xn_hi{s->Wnt, m, Kp, ldw, true}, Xn_lo{s->Wnl, m, Kp, ldw, true};
      s->gemmPosT.resize(T);
      for (int t = 1; t < T; ++t) {
        const size_t off = static_cast<size_t>(t) * K * KTp;
        MatRef Gt_hi{s->gramH.gtf + off, Kp, K, KTp, false}, Gt_lo{s->gramH.glo + off, Kp, K, KTp, false};
        ExtraSegs eP = three(Xn_hi, Xn_lo, Gt_hi, Gt_lo);
        NMFB_TRY(plan_store(h, ar, &s->gemmPosT[t], Xn_hi, Gt_hi, Kp, nullptr, nullptr, 0, m, Kp, s->Bt, nullptr, ldw,
                            false, nullptr, &eP));
        s->gemmPosT[t].L.args.chunk_kb = kChunk;
      }
      for (GemmOp* op : {&s->gemmC, &s->gemmD1}) op->L.args.chunk_kb = kChunk;
    }
  }

  std::vector<double> cost(static_cast<size_t>(cfg.maxiter) + 1, 0.0);  // cnmfsc.m:151
  if (spW) {  // cnmfsc.m:152-153: V_hat = RFD(W, H) with the projected W
    const bool keep = s->fused_resid;
    NMFB_TRY(objective(h, s, s->residInit, s->rsInit, &cost[0]));
    s->fused_resid = keep;
  } else {
    NMFB_TRY(objective(h, s, s->residCur, s->rsCur, &cost[0]));  // W == W0 here
  }
  std::vector<double> stepW(static_cast<size_t>(T), 1.0);  // cnmfsc.m:147
  double stepH = 1.0;                                                    // cnmfsc.m:148
  int ncost = cfg.maxiter + 1;
  bool done = false;
  for (int it = 1; it <= cfg.maxiter && !done; ++it) {
    if (!H_fixed) {
      NMFB_TRY(run_gemm(h, s->gemmN));
      if (spW && it == 1) {  // V_hat still is RFD(W, H) with the projected W (cnmfsc.m:152,164)
        NMFB_TRY(run_gemm(h, s->gemmC));
        NMFB_TRY(split_to(h, s->Cx, s->Cxt, s->Cxl, KTp, KTp, KTp));
        NMFB_TRY(run_gemm(h, s->gemmD1));
      } else {
        NMFB_TRY(run_gram(h, s->gramW, nullptr));
        NMFB_TRY(run_gemm(h, s->gemmD));
      }
      fold_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->P, s->Nf, K, T, n, ldh);
      NMFB_TRY(check_launch(h, "fold(N)"));
      fold_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->D, s->Df, K, T, n, ldh);
      NMFB_TRY(check_launch(h, "fold(D)"));
      if (sH > 0) {
        const double begobj = cost[it - 1];  // cnmfsc.m:169
        while (true) {
          grad_step_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, s->Df, s->Nf, s->Hnew, K, n, ldh, stepH);
          NMFB_TRY(check_launch(h, "grad_step(H)"));
          NMFB_TRY(project_rows(h, s, s->Hnew, L1s));  // cnmfsc.m:175-177
          NMFB_TRY(stack_split(h, s, s->Hnew, s->Hn, s->Hnt, s->Hnl));
          double newobj;
          NMFB_TRY(objective(h, s, s->residTrial, s->rsTrial, &newobj));  // cnmfsc.m:180-181
          if (newobj <= begobj) break;                                    // cnmfsc.m:184-186
          stepH /= 2;                                                     // cnmfsc.m:189
          if (stepH < 1e-200) {                                           // cnmfsc.m:190-194
            ncost = it;
            done = true;
            break;
          }
        }
        if (done) break;
        stepH *= 1.2;  // cnmfsc.m:198
        NMFB_CUDA(h, cudaMemcpyAsync(s->Hm, s->Hnew, ch * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        NMFB_CUDA(h, cudaMemcpyAsync(s->Hst, s->Hnt, chs * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        NMFB_CUDA(h, cudaMemcpyAsync(s->Hsl, s->Hnl, chs * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
      } else {  // cnmfsc.m:202-209
        mu_step_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, s->Nf, s->Df, n, ldh);
        NMFB_TRY(check_launch(h, "mu_step(H)"));
        NMFB_CUDA(h, cudaMemsetAsync(s->sq, 0, Kp * sizeof(double), h->stream));
        vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, ldh, nullptr, s->sq, nullptr);
        NMFB_TRY(check_launch(h, "vec_sums(H)"));
        for (int t = 0; t < T; ++t) {  // frame 0 also normalises the rows of H; the others only scale W
          renorm_pair_kernel<<<vec_grid(std::max(m, n), K), 256, 0, h->stream>>>(
              s->Hm, t == 0 ? n : 0, ldh, s->Wm + static_cast<size_t>(t) * K * ldw, m, ldw, s->sq);
          NMFB_TRY(check_launch(h, "renorm_pair"));
        }
        NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, KT, m, ldw));
        NMFB_TRY(stack_split(h, s, s->Hm, s->Hs, s->Hst, s->Hsl));
      }
    }
    if (!W_fixed && spW) {  // cnmfsc.m:216-254
      NMFB_TRY(run_gram(h, s->gramH, nullptr));
      NMFB_TRY(run_gemm(h, s->gemmA));
      double begobj;
      NMFB_TRY(objective(h, s, s->residCur, s->rsCur, &begobj));  // cnmfsc.m:215,218 (frame 1: the full model)
      for (int t = 0; t < T && !done; ++t) {
        const size_t off = static_cast<size_t>(t) * K * ldw;
        // pos = V_hat Hs_t' with the V_hat left by the previous frame's accepted trial (line 222)
        NMFB_TRY(run_gemm(h, t == 0 ? s->gemmBt[0] : s->gemmPosT[t]));
        double newobj = 0.0;
        while (true) {
          grad_step_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm + off, s->Bt, s->A + off, s->Wnew, K, m, ldw,
                                                                  stepW[t]);  // cnmfsc.m:229
          NMFB_TRY(check_launch(h, "grad_step(W)"));
          launch_projfunc(h->stream, K, s->Wnew, m, ldw, L1a, 1.0, 1, nullptr, s->fail);  // 230-232
          NMFB_TRY(check_launch(h, "projfunc(W)"));
          NMFB_TRY(split_to(h, s->Wnew, s->Wnt, s->Wnl, K, m, ldw));
          const bool keep = s->fused_resid;
          s->fused_resid = s->fused_w;
          int rc = objective(h, s, s->residW, s->rsW, &newobj);  // cnmfsc.m:235-236
          s->fused_resid = keep;
          NMFB_TRY(rc);
          if (newobj <= begobj) break;  // cnmfsc.m:239-241
          stepW[t] /= 2;                // cnmfsc.m:244
          if (stepW[t] < 1e-200) {      // cnmfsc.m:245-249
            ncost = it;
            done = true;
            break;
          }
        }
        if (done) break;
        stepW[t] *= 1.2;  // cnmfsc.m:252
        NMFB_CUDA(h, cudaMemcpyAsync(s->Wret + off, s->Wnew, static_cast<size_t>(K) * ldw * sizeof(float),
                                     cudaMemcpyDeviceToDevice, h->stream));  // cnmfsc.m:253
        begobj = newobj;  // line 218 of the next frame sees the V_hat of this trial
      }
      if (done) break;
      NMFB_CUDA(h, cudaMemcpyAsync(s->Wm, s->Wret, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));  // 266
      NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, KT, m, ldw));
    } else if (!W_fixed) {  // cnmfsc.m:257-263
      NMFB_TRY(run_gram(h, s->gramH, nullptr));
      NMFB_TRY(run_gemm(h, s->gemmA));
      for (int t = 0; t < T; ++t) {
        NMFB_TRY(run_gemm(h, s->gemmBt[t]));
        const size_t off = static_cast<size_t>(t) * K * ldw;
        mu_step_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm + off, s->A + off, s->Bt, m, ldw);
        NMFB_TRY(check_launch(h, "mu_step(W)"));
        NMFB_TRY(split_to(h, s->Wm + off, s->Wt + off, s->Wl + off, K, m, ldw));
      }
    } else if (spW || (!H_fixed && !(sH > 0))) {
      // cnmfsc.m:266 with W never updated: "W0 = W" replaces the unprojected W0 of the first iteration
      // and discards the scaling of lines 207-209 BEFORE the cost of this iteration is taken (H stays
      // normalised) - quirks of the reference, kept
      NMFB_CUDA(h, cudaMemcpyAsync(s->Wm, s->Wsave, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
      NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, KT, m, ldw));
    }
    NMFB_TRY(objective(h, s, s->residCur, s->rsCur, &cost[it]));  // cnmfsc.m:269-270
    if (it > 1 && cost[it] < cost[it - 1] && cost[it - 1] - cost[it] < cfg.tolerance) {  // cnmfsc.m:273-276
      ncost = it + 1;
      done = true;
    }
  }
  if (n_cost) *n_cost = ncost;
  if (cost_out) std::memcpy(cost_out, cost.data(), static_cast<size_t>(ncost) * sizeof(double));
  // the function returns W, which differs from W0 when it stops inside an iteration (cnmfsc.m:93-94,253,266)
  if (W_out) NMFB_TRY(download_colmajor(h, (spW && !W_fixed) ? s->Wret : (W_fixed ? s->Wsave : s->Wm), ldw, m, KT, W_out));
  if (H_out) NMFB_TRY(download_H(h, s->Hm, ldh, K, n, H_out));
  return NMFB_OK;
}

}  // namespace cnmfscdetail

extern "C" int nmfb_cnmfsc(nmfb_handle* h, int K, int T, const nmfb_config* cfg, float* W_out, float* H_out,
                           double* cost_out, int* n_cost) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  cnmfscdetail::State st;
  int rc = cnmfscdetail::run(h, &st, K, T, cfg, W_out, H_out, cost_out, n_cost);
  cudaStreamSynchronize(h->stream);
  return rc;
}
