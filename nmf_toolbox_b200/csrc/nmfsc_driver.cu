// nmfsc driver: Hoyer's NMF with sparseness constraints, nmfsc.m:57-245.
//
// H step then W step per iteration.  A constrained factor takes a projected
// gradient step (nmfsc.m:146-179 / 196-229): gradient from the tensor-core
// contractions, projfunc on every row / column, and the objective
// 0.5*|V - W*H|^2 of each trial evaluated explicitly by a fused
// reconstruct-and-residual kernel; the accept / halve decision of the line
// search is the reference's (newobj <= begobj) on those values.  An
// unconstrained factor takes the plain multiplicative step (182-188 / 232).
//
// Precision.  The projected-gradient step subtracts two nearly equal products
// (dH = W'V_hat - W'V) and the line search compares objectives that differ by
// parts in 1e5, so plain tf32 contractions (relative error ~1e-4 per operand)
// change the trajectory.  Every contraction here is therefore a split-tf32
// ("3xTF32") product: each operand is a tf32 head + tf32 tail and the tensor
// cores accumulate hi*hi + lo*hi + hi*lo in one pass (three operand segments),
// with short TMEM accumulation chunks; the result is fp32-accurate.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <ctime>
#include <vector>

#include "comm.cuh"
#include "engine.cuh"
#include "ew_kernels.cuh"

using namespace nmfb;

namespace nmfscdetail {

struct State {
  Arena ar;
  int K = 0, Kp = 0, m = 0, n = 0;
  long long ldw = 0, ldh = 0;
  // masters + tf32 head (t) / tail (l) pairs; "n" = line-search trial
  float *Wm, *Wt, *Wl, *Hm, *Ht, *Hl, *Wnew, *Wnt, *Wnl, *Hnew, *Hnt, *Hnl;
  float *Vhi, *Vlo;
  float *N, *D, *A, *B;
  double *scal, *sq;
  int* fail;
  GramOp gramW, gramH;
  GemmOp gemmN, gemmD, gemmA, gemmB, residCur, residH, residW;
  ResidOp rsCur, rsH, rsW;  // streaming objective kernel (resid_fused.cuh), K <= 128
  bool fused_resid = false;
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
  *out = 0.5 * v[0];  // nmfsc.m:139,161,212,238
  return NMFB_OK;
}

int split_to(nmfb_handle* h, const float* src, float* hi, float* lo, int nvec, int len, long long ld) {
  dim3 grid = vec_grid(len, nvec);
  grid.y = std::min<unsigned>(grid.y, 8192u);
  split_copy_kernel<<<grid, 256, 0, h->stream>>>(src, hi, lo, nvec, len, ld);
  return check_launch(h, "split_copy");
}

constexpr int kChunk = 2;  // k-blocks per TMEM accumulation chunk (8 MMA steps)

int project(nmfb_handle* h, State* s, float* X, int nvec, int len, long long ld, double k1) {
  if (len > kProjThreads * 32 * kProjMaskWords)
    return h->fail(NMFB_ERR_UNSUPPORTED, "nmfsc: projfunc vectors longer than %d are not supported",
                   kProjThreads * 32 * kProjMaskWords);
  projfunc_kernel<<<nvec, kProjThreads, 0, h->stream>>>(X, len, ld, k1, 1.0, 1, nullptr, s->fail);
  return check_launch(h, "projfunc");
}

int run(nmfb_handle* h, State* s, int K, const nmfb_config* cfg_in, float* W_out, float* H_out,
        double* cost_out, int* n_cost) {
  if (h->Vraw == nullptr) return h->fail(NMFB_ERR_NO_DATA, "nmfsc: call nmfb_set_V first");
  if (K <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "nmfsc: num_basis_elems must be positive");
  if (comm_size(h->comm) > 1) return h->fail(NMFB_ERR_UNSUPPORTED, "nmfsc: single GPU only");
  nmfb_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  if (cfg_in) cfg = *cfg_in;
  if (cfg.maxiter <= 0) cfg.maxiter = 100;         // nmfsc.m:123-125
  if (!(cfg.tolerance > 0)) cfg.tolerance = 1e-3;  // nmfsc.m:128-130
  double sW = cfg.W_sparsity, sH = cfg.H_sparsity;
  if (sW > 1) sW = 1;  // nmfsc.m:90-92
  if (sH > 1) sH = 1;  // nmfsc.m:103-105
  const bool W_fixed = cfg.W_fixed != 0, H_fixed = cfg.H_fixed != 0;
  const int m = h->m, n = h->n;
  s->K = K;
  s->Kp = round_up(K, 32);
  s->m = m;
  s->n = n;
  s->ldw = round_up(m, 4);
  s->ldh = round_up(n, 4);
  const int Kp = s->Kp;
  const long long ldw = s->ldw, ldh = s->ldh;
  Arena* ar = &s->ar;

  // nmfsc.m:57-62: reject negative data, rescale by the maximum
  VStats st;
  unsigned int* maxbits = nullptr;
  NMFB_TRY(compute_v_stats(h, false, &st, nullptr, &maxbits, ar));
  if (st.any_negative) return h->fail(NMFB_ERR_NEGATIVE_DATA, "Negative values in data!");
  NMFB_TRY(prepare_v_work(h, true, false, nullptr, maxbits));  // Vwork = V / max(V), unrounded
  NMFB_TRY(ar->alloc(h, &s->Vhi, static_cast<size_t>(n) * h->ldv));
  NMFB_TRY(ar->alloc(h, &s->Vlo, static_cast<size_t>(n) * h->ldv));
  NMFB_TRY(split_to(h, h->Vwork, s->Vhi, s->Vlo, n, static_cast<int>(h->ldv), h->ldv));

  const size_t cw = static_cast<size_t>(Kp) * ldw, ch = static_cast<size_t>(Kp) * ldh;
  NMFB_TRY(ar->alloc(h, &s->Wm, cw));
  NMFB_TRY(ar->alloc(h, &s->Wt, cw));
  NMFB_TRY(ar->alloc(h, &s->Wl, cw));
  NMFB_TRY(ar->alloc(h, &s->Wnew, cw));
  NMFB_TRY(ar->alloc(h, &s->Wnt, cw));
  NMFB_TRY(ar->alloc(h, &s->Wnl, cw));
  NMFB_TRY(ar->alloc(h, &s->A, cw));
  NMFB_TRY(ar->alloc(h, &s->B, cw));
  NMFB_TRY(ar->alloc(h, &s->Hm, ch));
  NMFB_TRY(ar->alloc(h, &s->Ht, ch));
  NMFB_TRY(ar->alloc(h, &s->Hl, ch));
  NMFB_TRY(ar->alloc(h, &s->Hnew, ch));
  NMFB_TRY(ar->alloc(h, &s->Hnt, ch));
  NMFB_TRY(ar->alloc(h, &s->Hnl, ch));
  NMFB_TRY(ar->alloc(h, &s->N, ch));
  NMFB_TRY(ar->alloc(h, &s->D, ch));
  NMFB_TRY(ar->alloc(h, &s->scal, 2));
  NMFB_TRY(ar->alloc(h, &s->sq, Kp));
  NMFB_TRY(ar->alloc(h, &s->fail, 1));

  {  // nmfsc.m:73-84
    std::vector<float> tmp;
    const float* Wsrc = cfg.W_init;
    if (!Wsrc) {
      tmp.resize(static_cast<size_t>(m) * K);
      fill_uniform(tmp, cfg.seed * 2 + 1, false);
      Wsrc = tmp.data();
    }
    NMFB_TRY(upload_colmajor(h, Wsrc, m, K, s->Wm, ldw));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    const float* Hsrc = cfg.H_init;
    if (!Hsrc) {
      tmp.resize(static_cast<size_t>(K) * n);
      fill_uniform(tmp, cfg.seed * 2 + 2, false);
      Hsrc = tmp.data();
    }
    NMFB_TRY(upload_H(h, ar, Hsrc, K, n, s->Hm, ldh));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!cfg.H_init) {  // default H_init has unit-L2 rows (nmfsc.m:80)
      vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, ldh, nullptr, s->sq, nullptr);
      NMFB_TRY(check_launch(h, "vec_sums"));
      renorm_pair_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, n, ldh, s->Wm, 0, ldw, s->sq);
      NMFB_TRY(check_launch(h, "renorm"));
    }
  }
  double L1a = 0, L1s = 0;
  if (sW > 0) {  // nmfsc.m:93-96
    L1a = std::sqrt(static_cast<double>(m)) - (std::sqrt(static_cast<double>(m)) - 1) * sW;
    NMFB_TRY(project(h, s, s->Wm, K, m, ldw, L1a));
  }
  if (sH > 0) {  // nmfsc.m:106-109
    L1s = std::sqrt(static_cast<double>(n)) - (std::sqrt(static_cast<double>(n)) - 1) * sH;
    NMFB_TRY(project(h, s, s->Hm, K, n, ldh, L1s));
  }
  NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, K, m, ldw));
  NMFB_TRY(split_to(h, s->Hm, s->Ht, s->Hl, K, n, ldh));

  // ---- contractions
  NMFB_TRY(plan_gram(h, ar, &s->gramW, s->Wt, Kp, m, ldw, nullptr, s->Wl, kChunk));
  NMFB_TRY(plan_gram(h, ar, &s->gramH, s->Ht, Kp, n, ldh, nullptr, s->Hl, kChunk));
  {
    auto three = [](const MatRef& Xhi, const MatRef& Xlo, const MatRef& Yhi, const MatRef& Ylo) {
      ExtraSegs e;  // acc = Xhi*Yhi' (segment 0) + Xlo*Yhi' + Xhi*Ylo'
      e.n = 2;
      e.X[0] = Xlo;
      e.Y[0] = Yhi;
      e.X[1] = Xhi;
      e.Y[1] = Ylo;
      return e;
    };
    // N = W'V (nmfsc.m:144): rows = columns j of V (K-major), contraction over i
    MatRef Vk_hi{s->Vhi, m, n, h->ldv, false}, Vk_lo{s->Vlo, m, n, h->ldv, false};
    MatRef Wk_hi{s->Wt, m, Kp, ldw, false}, Wk_lo{s->Wl, m, Kp, ldw, false};
    ExtraSegs eN = three(Vk_hi, Vk_lo, Wk_hi, Wk_lo);
    const int tiles_h = (n + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    NMFB_TRY(plan_store(h, ar, &s->gemmN, Vk_hi, Wk_hi, m, nullptr, nullptr, 0, n, Kp, s->N, nullptr, ldh,
                        tiles_h * 2 <= h->num_sms, nullptr, &eN));
    // D = W'V_hat = (W'W) H (nmfsc.m:145): rows j (H is MN-major there), contraction over k
    MatRef Hm_hi{s->Ht, n, Kp, ldh, true}, Hm_lo{s->Hl, n, Kp, ldh, true};
    MatRef Gw_hi{s->gramW.gtf, Kp, Kp, Kp, false}, Gw_lo{s->gramW.glo, Kp, Kp, Kp, false};
    ExtraSegs eD = three(Hm_hi, Hm_lo, Gw_hi, Gw_lo);
    NMFB_TRY(plan_store(h, ar, &s->gemmD, Hm_hi, Gw_hi, Kp, nullptr, nullptr, 0, n, Kp, s->D, nullptr, ldh,
                        false, nullptr, &eD));
    // A = V H' (nmfsc.m:194): rows i of V (MN-major), contraction over j
    MatRef Vm_hi{s->Vhi, m, n, h->ldv, true}, Vm_lo{s->Vlo, m, n, h->ldv, true};
    MatRef Hk_hi{s->Ht, n, Kp, ldh, false}, Hk_lo{s->Hl, n, Kp, ldh, false};
    ExtraSegs eA = three(Vm_hi, Vm_lo, Hk_hi, Hk_lo);
    const int tiles_w = (m + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    NMFB_TRY(plan_store(h, ar, &s->gemmA, Vm_hi, Hk_hi, n, nullptr, nullptr, 0, m, Kp, s->A, nullptr, ldw,
                        tiles_w * 2 <= h->num_sms, nullptr, &eA));
    // B = V_hat H' = W (H H') (nmfsc.m:195)
    MatRef Wm_hi{s->Wt, m, Kp, ldw, true}, Wm_lo{s->Wl, m, Kp, ldw, true};
    MatRef Gh_hi{s->gramH.gtf, Kp, Kp, Kp, false}, Gh_lo{s->gramH.glo, Kp, Kp, Kp, false};
    ExtraSegs eB = three(Wm_hi, Wm_lo, Gh_hi, Gh_lo);
    NMFB_TRY(plan_store(h, ar, &s->gemmB, Wm_hi, Gh_hi, Kp, nullptr, nullptr, 0, m, Kp, s->B, nullptr, ldw,
                        false, nullptr, &eB));
    for (GemmOp* op : {&s->gemmN, &s->gemmD, &s->gemmA, &s->gemmB}) op->L.args.chunk_kb = kChunk;
    // objective 0.5*|V - W*H|^2 of (W, H) pairs given as head/tail
    auto plan_resid = [&](GemmOp* op, const float* Whi, const float* Wlo, const float* Hhi, const float* Hlo) {
      MatRef Xh{Whi, m, Kp, ldw, true}, Xl{Wlo, m, Kp, ldw, true};
      MatRef Yh{Hhi, n, Kp, ldh, true}, Yl{Hlo, n, Kp, ldh, true};
      ExtraSegs e = three(Xh, Xl, Yh, Yl);
      NMFB_TRY(plan_fused(h, op, EPI_RESID, Xh, Yh, Kp, nullptr, nullptr, 0, m, round_up(n, 64), n, nullptr, &e));
      op->L.args.Vsrc = h->Vwork;
      op->L.args.ldv = h->ldv;
      op->L.args.scal = s->scal;
      op->L.args.chunk_kb = kChunk;
      std::string pe = set_v_prefetch(&op->L, h->Vwork, m, n, h->ldv);
      if (!pe.empty()) return h->fail(NMFB_ERR_CUDA, "%s", pe.c_str());
      return static_cast<int>(NMFB_OK);
    };
    s->fused_resid = Kp <= kKlMaxKp && std::getenv("NMFB_RESID_UNFUSED") == nullptr;
    if (s->fused_resid) {
      NMFB_TRY(nmfb::plan_resid(h, &s->rsCur, s->Wt, s->Wl, ldw, s->Ht, s->Hl, ldh, h->Vwork, h->ldv, m, n, Kp, s->scal));
      NMFB_TRY(nmfb::plan_resid(h, &s->rsH, s->Wt, s->Wl, ldw, s->Hnt, s->Hnl, ldh, h->Vwork, h->ldv, m, n, Kp, s->scal));
      NMFB_TRY(nmfb::plan_resid(h, &s->rsW, s->Wnt, s->Wnl, ldw, s->Ht, s->Hl, ldh, h->Vwork, h->ldv, m, n, Kp, s->scal));
    } else {
      NMFB_TRY(plan_resid(&s->residCur, s->Wt, s->Wl, s->Ht, s->Hl));
      NMFB_TRY(plan_resid(&s->residH, s->Wt, s->Wl, s->Hnt, s->Hnl));
      NMFB_TRY(plan_resid(&s->residW, s->Wnt, s->Wnl, s->Ht, s->Hl));
    }
  }

  const bool trace = std::getenv("NMFB_TRACE") != nullptr;
  double tr[6] = {0, 0, 0, 0, 0, 0};
  int ntrials = 0;
  auto now = [] {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  auto lap = [&](int slot, double& t0) {
    if (!trace) return;
    cudaStreamSynchronize(h->stream);
    const double t1 = now();
    tr[slot] += t1 - t0;
    t0 = t1;
  };
  std::vector<double> cost(static_cast<size_t>(cfg.maxiter) + 1, 0.0);  // nmfsc.m:137
  NMFB_TRY(objective(h, s, s->residCur, s->rsCur, &cost[0]));                      // nmfsc.m:138-139
  double stepW = 1.0, stepH = 1.0;                                       // nmfsc.m:133-134
  int ncost = cfg.maxiter + 1;
  bool done = false;
  for (int it = 1; it <= cfg.maxiter && !done; ++it) {
    double tl = now();
    if (!H_fixed) {
      NMFB_TRY(run_gram(h, s->gramW, nullptr));
      NMFB_TRY(run_gemm(h, s->gemmN));  // N = W'V (144)
      NMFB_TRY(run_gemm(h, s->gemmD));  // D = W'V_hat = (W'W)H (145)
      lap(0, tl);
      if (sH > 0) {
        const double begobj = cost[it - 1];  // nmfsc.m:149
        while (true) {
          grad_step_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, s->D, s->N, s->Hnew, K, n, ldh, stepH);
          NMFB_TRY(check_launch(h, "grad_step(H)"));
          NMFB_TRY(project(h, s, s->Hnew, K, n, ldh, L1s));  // nmfsc.m:155-157
          NMFB_TRY(split_to(h, s->Hnew, s->Hnt, s->Hnl, K, n, ldh));
          lap(1, tl);
          ++ntrials;
          double newobj;
          NMFB_TRY(objective(h, s, s->residH, s->rsH, &newobj));  // nmfsc.m:160-161
          lap(2, tl);
          if (newobj <= begobj) break;                    // nmfsc.m:164-166
          stepH /= 2;                                     // nmfsc.m:169
          if (stepH < 1e-200) {                           // nmfsc.m:170-174
            ncost = it;
            done = true;
            break;
          }
        }
        if (done) break;
        stepH *= 1.2;  // nmfsc.m:178
        NMFB_CUDA(h, cudaMemcpyAsync(s->Hm, s->Hnew, ch * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        NMFB_CUDA(h, cudaMemcpyAsync(s->Ht, s->Hnt, ch * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        NMFB_CUDA(h, cudaMemcpyAsync(s->Hl, s->Hnl, ch * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
      } else {  // nmfsc.m:182-187
        mu_step_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, s->N, s->D, n, ldh);
        NMFB_TRY(check_launch(h, "mu_step(H)"));
        NMFB_CUDA(h, cudaMemsetAsync(s->sq, 0, Kp * sizeof(double), h->stream));
        vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, ldh, nullptr, s->sq, nullptr);
        NMFB_TRY(check_launch(h, "vec_sums(H)"));
        renorm_pair_kernel<<<vec_grid(std::max(m, n), K), 256, 0, h->stream>>>(s->Hm, n, ldh, s->Wm, m, ldw, s->sq);
        NMFB_TRY(check_launch(h, "renorm_pair"));
        NMFB_TRY(split_to(h, s->Hm, s->Ht, s->Hl, K, n, ldh));
        NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, K, m, ldw));
      }
    }
    lap(3, tl);
    if (!W_fixed) {
      NMFB_TRY(run_gram(h, s->gramH, nullptr));
      NMFB_TRY(run_gemm(h, s->gemmA));  // A = V H' (194)
      NMFB_TRY(run_gemm(h, s->gemmB));  // B = V_hat H' = W (H H') (195)
      if (sW > 0) {
        double begobj;
        NMFB_TRY(objective(h, s, s->residCur, s->rsCur, &begobj));  // nmfsc.m:193,197
        while (true) {
          grad_step_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm, s->B, s->A, s->Wnew, K, m, ldw, stepW);
          NMFB_TRY(check_launch(h, "grad_step(W)"));
          NMFB_TRY(project(h, s, s->Wnew, K, m, ldw, L1a));  // nmfsc.m:206-208
          NMFB_TRY(split_to(h, s->Wnew, s->Wnt, s->Wnl, K, m, ldw));
          double newobj;
          NMFB_TRY(objective(h, s, s->residW, s->rsW, &newobj));  // nmfsc.m:211-212
          if (newobj <= begobj) break;
          stepW /= 2;
          if (stepW < 1e-200) {  // nmfsc.m:221-225
            ncost = it;
            done = true;
            break;
          }
        }
        if (done) break;
        stepW *= 1.2;
        NMFB_CUDA(h, cudaMemcpyAsync(s->Wm, s->Wnew, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        NMFB_CUDA(h, cudaMemcpyAsync(s->Wt, s->Wnt, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        NMFB_CUDA(h, cudaMemcpyAsync(s->Wl, s->Wnl, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
      } else {  // nmfsc.m:232
        mu_step_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm, s->A, s->B, m, ldw);
        NMFB_TRY(check_launch(h, "mu_step(W)"));
        NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, K, m, ldw));
      }
    }
    lap(4, tl);
    NMFB_TRY(objective(h, s, s->residCur, s->rsCur, &cost[it]));  // nmfsc.m:237-238
    lap(5, tl);
    if (it > 1 && cost[it] < cost[it - 1] && cost[it - 1] - cost[it] < cfg.tolerance) {  // 241-244
      ncost = it + 1;
      done = true;
    }
  }
  if (trace)
    fprintf(stderr, "[nmfb] nmfsc per-phase ms over the run: gradient GEMMs %.2f, trial prep (step+projfunc+split) %.2f, "
                    "trial objective %.2f (%d trials), copies %.2f, W step %.2f, final objective %.2f\n",
            tr[0], tr[1], tr[2], ntrials, tr[3], tr[4], tr[5]);
  if (n_cost) *n_cost = ncost;
  if (cost_out) std::memcpy(cost_out, cost.data(), static_cast<size_t>(ncost) * sizeof(double));
  if (W_out) NMFB_TRY(download_colmajor(h, s->Wm, ldw, m, K, W_out));
  if (H_out) NMFB_TRY(download_H(h, s->Hm, ldh, K, n, H_out));
  return NMFB_OK;
}

}  // namespace nmfscdetail

extern "C" int nmfb_nmfsc(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out,
                          double* cost_out, int* n_cost) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  nmfscdetail::State st;
  int rc = nmfscdetail::run(h, &st, K, cfg, W_out, H_out, cost_out, n_cost);
  cudaStreamSynchronize(h->stream);
  return rc;
}
