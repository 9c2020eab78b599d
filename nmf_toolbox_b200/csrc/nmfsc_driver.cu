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
// The line searches run on the device (ew_kernels.cuh, "device-side line search"): the host queues a
// fixed pattern of phase-guarded kernels per iteration and never waits for an accept / halve decision.
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
  double *scal, *sq, *cost;
  int* fail;
  unsigned int* ticket;
  LsState* ls;
  GramOp gramW, gramH;
  GemmOp gemmN, gemmD, gemmA, gemmB, residCur, residH, residW;
  ResidOp rsCur, rsH, rsW;  // streaming objective kernel (resid_fused.cuh), K <= 128
  bool fused_resid = false;
};

// Objective 0.5*|V - W*H|^2 (x2 in scal[0]) guarded by the phase word `skip`, followed by what the line
// search does with it (`fin`: decide a trial, close the iteration, ...).  The streaming kernel lets its last
// block do that; the panel-GEMM fallback (K > 128) needs the small single-thread kernel.
int objective(nmfb_handle* h, State* s, GemmOp& op, ResidOp& rs, const int* skip, int fin) {
  NMFB_TRY(prof_mark(h, 1));
  if (s->fused_resid) {
    rs.args.skip = skip;
    rs.args.fin.mode = fin;
    rs.args.fin.st = s->ls;
    rs.args.fin.cost = s->cost;
    rs.args.fin.fail = s->fail;
    rs.args.fin.ticket = s->ticket;
    NMFB_TRY(run_resid(h, rs));
  } else {
    op.L.args.stop = skip;
    NMFB_TRY(run_gemm(h, op));
  }
  NMFB_TRY(prof_mark(h, 1));
  if (s->fused_resid || fin == LSFIN_NONE) return NMFB_OK;
  if (fin == LSFIN_DECIDE_H || fin == LSFIN_DECIDE_W) {
    ls_decide_kernel<<<1, 1, 0, h->stream>>>(s->ls, fin == LSFIN_DECIDE_W ? 1 : 0, s->scal, s->fail);
    return check_launch(h, "ls_decide");
  }
  if (fin == LSFIN_COST) {
    ls_cost_kernel<<<1, 1, 0, h->stream>>>(s->ls, s->scal, s->cost);
    return check_launch(h, "ls_cost");
  }
  ls_advance_kernel<<<1, 1, 0, h->stream>>>(s->ls, fin == LSFIN_INIT ? LS_INIT : LS_TO_WTRIAL, s->scal, s->cost, s->fail);
  return check_launch(h, "ls_advance");
}

int split_to(nmfb_handle* h, const float* src, float* hi, float* lo, int nvec, int len, long long ld,
             const int* skip = nullptr) {
  dim3 grid = vec_grid(len, nvec);
  grid.y = std::min<unsigned>(grid.y, 8192u);
  split_copy_kernel<<<grid, 256, 0, h->stream>>>(src, hi, lo, nvec, len, ld, skip);
  return check_launch(h, "split_copy");
}

// k-blocks per TMEM accumulation chunk of the gradient contractions (x 4 MMA steps each): the tensor core
// truncates when it adds into the accumulator, so long chunks bias the sums; short ones cost drain time
static int sc_chunk() {
  static int v = [] {
    const char* e = std::getenv("NMFB_SC_CHUNK");
    const int c = e ? std::atoi(e) : 8;  // parity and the halving sequences are unchanged for 2..16 (measured)
    return c > 0 ? c : 8;
  }();
  return v;
}
static int sc_gram_ctas() {  // CTAs (= split-K slabs to sum afterwards) of the small Gram products
  static int v = [] {
    const char* e = std::getenv("NMFB_SC_GRAM_CTAS");
    const int c = e ? std::atoi(e) : 0;
    return c > 0 ? c : 0;
  }();
  return v;
}

// The iteration alternates tensor-core kernels that need ~200 KB of shared memory with small vector
// kernels that need almost none.  Left alone, the driver picks a different L1 / shared-memory split for
// the two kinds and re-partitions the SMs at every switch, which costs an idle gap per kernel boundary;
// asking for the maximal shared-memory carve-out on the small kernels keeps one configuration.
void prefer_max_shared_once() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* e = std::getenv("NMFB_NO_CARVEOUT");
  if (e && e[0] == '1') return;
  auto set = [](const void* fn) {
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  };
  set(reinterpret_cast<const void*>(projfunc_kernel_t<8>));
  set(reinterpret_cast<const void*>(projfunc_kernel_t<32>));
  set(reinterpret_cast<const void*>(projfunc_kernel_t<0>));
  set(reinterpret_cast<const void*>(ls_advance_kernel));
  set(reinterpret_cast<const void*>(ls_decide_kernel));
  set(reinterpret_cast<const void*>(ls_cost_kernel));
  set(reinterpret_cast<const void*>(copy3_kernel));
  set(reinterpret_cast<const void*>(mu_step_split_kernel));
  set(reinterpret_cast<const void*>(mu_step_kernel));
  set(reinterpret_cast<const void*>(split_copy_kernel));
  set(reinterpret_cast<const void*>(gram_reduce_kernel));
  set(reinterpret_cast<const void*>(split_reduce_kernel));
  set(reinterpret_cast<const void*>(vec_sums_kernel));
  set(reinterpret_cast<const void*>(renorm_pair_kernel));
  cudaGetLastError();
}

int project(nmfb_handle* h, State* s, float* X, int nvec, int len, long long ld, double k1,
            const ProjFuse& f = ProjFuse()) {
  if (len > kProjThreads * 32 * kProjMaskWords)
    return h->fail(NMFB_ERR_UNSUPPORTED, "nmfsc: projfunc vectors longer than %d are not supported",
                   kProjThreads * 32 * kProjMaskWords);
  launch_projfunc(h->stream, nvec, X, len, ld, k1, 1.0, 1, nullptr, s->fail, f);
  return check_launch(h, "projfunc");
}

int advance(nmfb_handle* h, State* s, int what) {
  ls_advance_kernel<<<1, 1, 0, h->stream>>>(s->ls, what, s->scal, s->cost, s->fail);
  return check_launch(h, "ls_advance");
}

int run(nmfb_handle* h, State* s, int K, const nmfb_config* cfg_in, float* W_out, float* H_out,
        double* cost_out, int* n_cost) {
  if (h->Vraw == nullptr) return h->fail(NMFB_ERR_NO_DATA, "nmfsc: call nmfb_set_V first");
  if (K <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "nmfsc: num_basis_elems must be positive");
  if (comm_size(h->comm) > 1) return h->fail(NMFB_ERR_UNSUPPORTED, "nmfsc: single GPU only");
  prefer_max_shared_once();
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
  NMFB_TRY(ar->alloc(h, &s->Wt, 2 * cw));  // tf32 head and tail back to back: one stacked operand [W_hi; W_lo]
  s->Wl = s->Wt + cw;
  NMFB_TRY(ar->alloc(h, &s->Wnew, cw));
  NMFB_TRY(ar->alloc(h, &s->Wnt, cw));
  NMFB_TRY(ar->alloc(h, &s->Wnl, cw));
  NMFB_TRY(ar->alloc(h, &s->A, cw));
  NMFB_TRY(ar->alloc(h, &s->B, cw));
  NMFB_TRY(ar->alloc(h, &s->Hm, ch));
  NMFB_TRY(ar->alloc(h, &s->Ht, 2 * ch));  // [H_hi; H_lo]
  s->Hl = s->Ht + ch;
  NMFB_TRY(ar->alloc(h, &s->Hnew, ch));
  NMFB_TRY(ar->alloc(h, &s->Hnt, ch));
  NMFB_TRY(ar->alloc(h, &s->Hnl, ch));
  NMFB_TRY(ar->alloc(h, &s->N, ch));
  NMFB_TRY(ar->alloc(h, &s->D, ch));
  NMFB_TRY(ar->alloc(h, &s->scal, 2));
  NMFB_TRY(ar->alloc(h, &s->sq, Kp));
  NMFB_TRY(ar->alloc(h, &s->fail, 1));
  NMFB_TRY(ar->alloc(h, &s->cost, static_cast<size_t>(cfg.maxiter) + 1));  // nmfsc.m:137
  NMFB_TRY(ar->alloc(h, &s->ls, 1));
  NMFB_TRY(ar->alloc(h, &s->ticket, 1));
  {
    LsState init{};
    init.stepH = init.stepW = 1.0;  // nmfsc.m:133-134
    init.skip[0] = 0;
    init.skip[1] = init.skip[2] = init.skip[3] = init.skip[4] = 1;
    init.maxiter = cfg.maxiter;
    init.tolerance = cfg.tolerance;
    NMFB_TRY(ar->alloc(h, &init.halvings, 2 * static_cast<size_t>(cfg.maxiter)));
    NMFB_CUDA(h, cudaMemcpyAsync(s->ls, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  const int* sk0 = &s->ls->skip[0];
  const int* sk1 = &s->ls->skip[1];
  const int* sk2 = &s->ls->skip[2];
  const int* sk3 = &s->ls->skip[3];
  const int* sk4 = &s->ls->skip[4];

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
  NMFB_TRY(plan_gram(h, ar, &s->gramW, s->Wt, Kp, m, ldw, sk0, s->Wl, sc_chunk(), sc_gram_ctas()));
  NMFB_TRY(plan_gram(h, ar, &s->gramH, s->Ht, Kp, n, ldh, sk2, s->Hl, sc_chunk(), sc_gram_ctas()));
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
    // The two contractions with V stream it as a tf32 head / tail pair (2 x m x n floats).  Written as three
    // segments (V_hi W_hi + V_lo W_hi + V_hi W_lo) the head travels twice; with the small factor STACKED as
    // [W_hi; W_lo] (twice the accumulator columns) two segments do, V_hi [W_hi; W_lo]' + V_lo [W_hi; W_lo]', and the
    // two column halves are added when the split-K slabs are summed (the lo*lo term comes for free).
    const bool folded = std::getenv("NMFB_SC_UNFOLDED") == nullptr;
    auto plan_folded = [&](GemmOp* op, const MatRef& Xhi, const MatRef& Xlo, const MatRef& Y2, long long kdim, int rows,
                           float* out, long long ldo, bool allow_split, const int* stop) -> int {
      ExtraSegs e;
      e.n = 1;
      e.X[0] = Xlo;
      e.Y[0] = Y2;
      float* whole = nullptr;  // [2 Kp][ldo] when the contraction is not split
      NMFB_TRY(ar->alloc(h, &whole, static_cast<size_t>(2 * Kp) * ldo));
      NMFB_TRY(plan_store(h, ar, op, Xhi, Y2, kdim, nullptr, nullptr, 0, rows, 2 * Kp, whole, nullptr, ldo, allow_split,
                          stop, &e));
      if (op->splits == 1) {
        op->parts = whole;
        op->L.args.out0 = whole;
        op->L.args.split_stride = 0;
      }
      op->splits *= 2;  // every slab is two half-slabs of Kp x ldo
      op->count = static_cast<long long>(Kp) * ldo;
      op->final0 = out;
      return static_cast<int>(NMFB_OK);
    };
    ExtraSegs eN = three(Vk_hi, Vk_lo, Wk_hi, Wk_lo);
    const int tiles_h = (n + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    if (folded) {
      MatRef Wk2{s->Wt, m, 2 * Kp, ldw, false};
      NMFB_TRY(plan_folded(&s->gemmN, Vk_hi, Vk_lo, Wk2, m, n, s->N, ldh, tiles_h * 2 <= h->num_sms, sk0));
    } else {
    NMFB_TRY(plan_store(h, ar, &s->gemmN, Vk_hi, Wk_hi, m, nullptr, nullptr, 0, n, Kp, s->N, nullptr, ldh,
                        tiles_h * 2 <= h->num_sms, sk0, &eN));
    }
    // D = W'V_hat = (W'W) H (nmfsc.m:145): rows j (H is MN-major there), contraction over k
    MatRef Hm_hi{s->Ht, n, Kp, ldh, true}, Hm_lo{s->Hl, n, Kp, ldh, true};
    MatRef Gw_hi{s->gramW.gtf, Kp, Kp, Kp, false}, Gw_lo{s->gramW.glo, Kp, Kp, Kp, false};
    ExtraSegs eD = three(Hm_hi, Hm_lo, Gw_hi, Gw_lo);
    NMFB_TRY(plan_store(h, ar, &s->gemmD, Hm_hi, Gw_hi, Kp, nullptr, nullptr, 0, n, Kp, s->D, nullptr, ldh,
                        false, sk0, &eD));
    // A = V H' (nmfsc.m:194): rows i of V (MN-major), contraction over j
    MatRef Vm_hi{s->Vhi, m, n, h->ldv, true}, Vm_lo{s->Vlo, m, n, h->ldv, true};
    MatRef Hk_hi{s->Ht, n, Kp, ldh, false}, Hk_lo{s->Hl, n, Kp, ldh, false};
    ExtraSegs eA = three(Vm_hi, Vm_lo, Hk_hi, Hk_lo);
    const int tiles_w = (m + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    if (folded) {
      MatRef Hk2{s->Ht, n, 2 * Kp, ldh, false};
      NMFB_TRY(plan_folded(&s->gemmA, Vm_hi, Vm_lo, Hk2, n, m, s->A, ldw, tiles_w * 2 <= h->num_sms, sk2));
    } else {
    NMFB_TRY(plan_store(h, ar, &s->gemmA, Vm_hi, Hk_hi, n, nullptr, nullptr, 0, m, Kp, s->A, nullptr, ldw,
                        tiles_w * 2 <= h->num_sms, sk2, &eA));
    }
    // B = V_hat H' = W (H H') (nmfsc.m:195)
    MatRef Wm_hi{s->Wt, m, Kp, ldw, true}, Wm_lo{s->Wl, m, Kp, ldw, true};
    MatRef Gh_hi{s->gramH.gtf, Kp, Kp, Kp, false}, Gh_lo{s->gramH.glo, Kp, Kp, Kp, false};
    ExtraSegs eB = three(Wm_hi, Wm_lo, Gh_hi, Gh_lo);
    NMFB_TRY(plan_store(h, ar, &s->gemmB, Wm_hi, Gh_hi, Kp, nullptr, nullptr, 0, m, Kp, s->B, nullptr, ldw,
                        false, sk2, &eB));
    for (GemmOp* op : {&s->gemmN, &s->gemmD, &s->gemmA, &s->gemmB}) op->L.args.chunk_kb = sc_chunk();
    // objective 0.5*|V - W*H|^2 of (W, H) pairs given as head/tail
    auto plan_resid = [&](GemmOp* op, const float* Whi, const float* Wlo, const float* Hhi, const float* Hlo) {
      MatRef Xh{Whi, m, Kp, ldw, true}, Xl{Wlo, m, Kp, ldw, true};
      MatRef Yh{Hhi, n, Kp, ldh, true}, Yl{Hlo, n, Kp, ldh, true};
      ExtraSegs e = three(Xh, Xl, Yh, Yl);
      NMFB_TRY(plan_fused(h, op, EPI_RESID, Xh, Yh, Kp, nullptr, nullptr, 0, m, round_up(n, 64), n, nullptr, &e));
      op->L.args.Vsrc = h->Vwork;
      op->L.args.ldv = h->ldv;
      op->L.args.scal = s->scal;
      op->L.args.chunk_kb = sc_chunk();
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
  const bool searchH = sH > 0 && !H_fixed, searchW = sW > 0 && !W_fixed;
  int slots = 2;  // line-search trial slots per pattern (a longer search continues in the next pattern)
  if (const char* e = std::getenv("NMFB_LS_SLOTS")) slots = std::max(1, std::atoi(e));
  const long long ch4 = static_cast<long long>(ch / 4), cw4 = static_cast<long long>(cw / 4);
  const int copy_blocks = h->num_sms * 4;

  // one trial slot of a line search (nmfsc.m:152-175 / 203-226): step + projfunc + split, objective, decision
  auto trial = [&](bool for_w) -> int {
    ProjFuse f;
    f.src = for_w ? s->Wm : s->Hm;
    f.Dp = for_w ? s->B : s->D;
    f.Dn = for_w ? s->A : s->N;
    f.step = for_w ? &s->ls->stepW : &s->ls->stepH;
    f.hi = for_w ? s->Wnt : s->Hnt;
    f.lo = for_w ? s->Wnl : s->Hnl;
    f.skip = for_w ? sk3 : sk1;
    NMFB_TRY(prof_mark(h, 2));
    if (for_w) NMFB_TRY(project(h, s, s->Wnew, K, m, ldw, L1a, f));  // nmfsc.m:205-208
    else NMFB_TRY(project(h, s, s->Hnew, K, n, ldh, L1s, f));        // nmfsc.m:154-157
    NMFB_TRY(prof_mark(h, 2));
    if (for_w) return objective(h, s, s->residW, s->rsW, sk3, LSFIN_DECIDE_W);  // nmfsc.m:211-212, 215-225
    return objective(h, s, s->residH, s->rsH, sk1, LSFIN_DECIDE_H);             // nmfsc.m:160-161, 164-174
  };
  // the kernels of one iteration (H first, then W: nmfsc.m:143-244), every one guarded by its phase
  auto pattern = [&]() -> int {
    // ---- phase 0: H gradient
    if (!H_fixed) {
      NMFB_TRY(prof_mark(h, 0));
      NMFB_TRY(run_gram(h, s->gramW, sk0));
      NMFB_TRY(run_gemm(h, s->gemmN));  // N = W'V (144)
      NMFB_TRY(run_gemm(h, s->gemmD));  // D = W'V_hat = (W'W)H (145)
      NMFB_TRY(prof_mark(h, 0));
      if (searchH) {
        NMFB_TRY(advance(h, s, LS_TO_HTRIAL));
      } else {  // nmfsc.m:182-187
        mu_step_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, s->N, s->D, n, ldh, sk0);
        NMFB_TRY(check_launch(h, "mu_step(H)"));
        NMFB_CUDA(h, cudaMemsetAsync(s->sq, 0, Kp * sizeof(double), h->stream));  // scratch of this phase only
        vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, ldh, nullptr, s->sq, sk0);
        NMFB_TRY(check_launch(h, "vec_sums(H)"));
        renorm_pair_kernel<<<vec_grid(std::max(m, n), K), 256, 0, h->stream>>>(s->Hm, n, ldh, s->Wm, m, ldw, s->sq, sk0);
        NMFB_TRY(check_launch(h, "renorm_pair"));
        NMFB_TRY(split_to(h, s->Hm, s->Ht, s->Hl, K, n, ldh, sk0));
        NMFB_TRY(split_to(h, s->Wm, s->Wt, s->Wl, K, m, ldw, sk0));
        NMFB_TRY(advance(h, s, LS_H_TO_WGRAD));
      }
    } else {
      NMFB_TRY(advance(h, s, LS_H_TO_WGRAD));
    }
    // ---- phase 1: H line search
    if (searchH)
      for (int t = 0; t < slots; ++t) NMFB_TRY(trial(false));
    // ---- phase 2: commit H (nmfsc.m:179), W gradient
    if (searchH) {
      copy3_kernel<<<copy_blocks, 256, 0, h->stream>>>(s->Hnew, s->Hm, s->Hnt, s->Ht, s->Hnl, s->Hl, ch4, sk2);
      NMFB_TRY(check_launch(h, "copy3(H)"));
    }
    if (!W_fixed) {
      NMFB_TRY(prof_mark(h, 3));
      NMFB_TRY(run_gram(h, s->gramH, sk2));
      NMFB_TRY(run_gemm(h, s->gemmA));  // A = V H' (194)
      NMFB_TRY(run_gemm(h, s->gemmB));  // B = V_hat H' = W (H H') (195)
      if (searchW) {
        NMFB_TRY(objective(h, s, s->residCur, s->rsCur, sk2, LSFIN_TO_WTRIAL));  // nmfsc.m:193,197
      } else {  // nmfsc.m:232
        mu_step_split_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm, s->A, s->B, s->Wt, s->Wl, m, ldw, sk2);
        NMFB_TRY(check_launch(h, "mu_step(W)"));
        NMFB_TRY(advance(h, s, LS_W_TO_COST));
      }
      NMFB_TRY(prof_mark(h, 3));
    } else {
      NMFB_TRY(advance(h, s, LS_W_TO_COST));
    }
    // ---- phase 3: W line search
    if (searchW)
      for (int t = 0; t < slots; ++t) NMFB_TRY(trial(true));
    // ---- phase 4: commit W (nmfsc.m:229), cost(iter+1) and stop test (237-244)
    if (searchW) {
      copy3_kernel<<<copy_blocks, 256, 0, h->stream>>>(s->Wnew, s->Wm, s->Wnt, s->Wt, s->Wnl, s->Wl, cw4, sk4);
      NMFB_TRY(check_launch(h, "copy3(W)"));
    }
    return objective(h, s, s->residCur, s->rsCur, sk4, LSFIN_COST);
  };

  NMFB_CUDA(h, cudaMemsetAsync(s->scal, 0, 2 * sizeof(double), h->stream));
  NMFB_TRY(objective(h, s, s->residCur, s->rsCur, nullptr, LSFIN_INIT));  // nmfsc.m:138-139
  loop_begin(h);
  // Queue patterns in chunks; the host looks at {iter, ncost, done, failed} of the chunk before the one
  // it has just queued, so it never waits for the device while work is outstanding.
  constexpr int kPatternsPerChunk = 8;
  int* pin = h->pinned + 8;  // two slots of four ints
  cudaEvent_t evs[2] = {nullptr, nullptr};
  for (int b = 0; b < 2; ++b)
    if (cudaEventCreateWithFlags(&evs[b], cudaEventDisableTiming) != cudaSuccess)
      return h->fail(NMFB_ERR_CUDA, "cudaEventCreate failed");
  int rc = NMFB_OK, c = 0, queued = 0;
  std::memset(pin, 0, 8 * sizeof(int));
  // The pattern is ~25 short kernels whose arguments never change (step sizes, objectives and the
  // iteration counter live in device memory), and launched one by one they are launch-latency bound
  // (measured: ~200 us of kernel time in a 360 us iteration).  So the first pattern is launched
  // directly (it also performs every one-off cudaFuncSetAttribute), the second is recorded into a CUDA
  // graph, and every further iteration is ONE graph launch.  Per-kernel event timing
  // (nmfb_profile_enable) and NMFB_NO_GRAPH=1 keep the direct launches.
  cudaGraphExec_t gexec = nullptr;
  long long per_pattern = 0;
  {
    const long long l0 = h->launches;
    rc = pattern();
    ++queued;
    per_pattern = h->launches - l0;
    const char* ng = std::getenv("NMFB_NO_GRAPH");
    if (rc == NMFB_OK && !h->profile && !(ng && ng[0] == '1')) {
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        const int prc = pattern();
        h->launches -= per_pattern;  // recorded, not launched
        cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
        if (prc != NMFB_OK) rc = prc;
        else if (e2 != cudaSuccess) rc = h->fail(NMFB_ERR_CUDA, "nmfsc: graph capture failed: %s", cudaGetErrorString(e2));
        else if ((e2 = cudaGraphInstantiate(&gexec, graph, 0)) != cudaSuccess)
          rc = h->fail(NMFB_ERR_CUDA, "nmfsc: cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
        if (graph) cudaGraphDestroy(graph);
      } else {
        rc = h->fail(NMFB_ERR_CUDA, "nmfsc: cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
      }
    }
  }
  auto queue_pattern = [&]() -> int {
    if (gexec == nullptr) return pattern();
    cudaError_t e = cudaGraphLaunch(gexec, h->stream);
    if (e != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "nmfsc: cudaGraphLaunch failed: %s", cudaGetErrorString(e));
    h->launches += per_pattern;
    return NMFB_OK;
  };
  // a search may halve ~665 times before the step underflows (nmfsc.m:170): generous bound, never reached
  const long long max_patterns = (static_cast<long long>(cfg.maxiter) + 2) * 700;
  while (rc == NMFB_OK) {
    for (int p = 0; p < kPatternsPerChunk && rc == NMFB_OK; ++p, ++queued) rc = queue_pattern();
    if (rc != NMFB_OK) break;
    if (queued > max_patterns) {
      rc = h->fail(NMFB_ERR_CUDA, "internal: nmfsc iteration loop did not finish after %d kernel patterns", queued);
      break;
    }
    if (c > 0) {
      cudaEventSynchronize(evs[(c - 1) & 1]);
      if (pin[((c - 1) & 1) * 4 + 2] != 0) break;  // done
    }
    cudaMemcpyAsync(pin + (c & 1) * 4, &s->ls->iter, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaEventRecord(evs[c & 1], h->stream);
    ++c;
  }
  for (int b = 0; b < 2; ++b) cudaEventDestroy(evs[b]);
  cudaStreamSynchronize(h->stream);
  if (gexec) cudaGraphExecDestroy(gexec);
  NMFB_TRY(rc);
  LsState fin;
  NMFB_CUDA(h, cudaMemcpy(&fin, s->ls, sizeof(fin), cudaMemcpyDeviceToHost));
  loop_end(h, fin.iter);
  if (fin.failed) return h->fail(NMFB_ERR_PROJFUNC, "projfunc diverged (non-finite values)");
  if (trace)
    fprintf(stderr, "[nmfb] nmfsc: %d iterations, %d line-search trials, %d patterns queued (%s), loop %.2f ms "
                    "(%.1f us per iteration), final steps H %.3g W %.3g\n",
            fin.iter, fin.trials, queued, gexec ? "CUDA graph" : "direct launches", h->loop_ms, fin.iter ? 1e3 * h->loop_ms / fin.iter : 0.0, fin.stepH, fin.stepW);
  h->halvings.assign(2 * static_cast<size_t>(fin.iter), 0);
  if (fin.iter > 0)
    NMFB_CUDA(h, cudaMemcpy(h->halvings.data(), fin.halvings, h->halvings.size() * sizeof(int), cudaMemcpyDeviceToHost));
  const int ncost = fin.ncost;
  if (n_cost) *n_cost = ncost;
  if (cost_out && ncost > 0)
    NMFB_CUDA(h, cudaMemcpy(cost_out, s->cost, static_cast<size_t>(ncost) * sizeof(double), cudaMemcpyDeviceToHost));
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
