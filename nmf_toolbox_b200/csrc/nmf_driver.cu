// nmf driver: the iteration loop of nmf.m (lines 143-225) for the Euclidean and
// KL divergences, as a queue of device kernels with no host synchronisation
// inside the loop.  Per iteration (nmf.m line numbers):
//
//   Euclidean                                       KL
//   G_H = H H'                                      hs = H * 1                       (153)
//   [cost of the previous iteration + stop test]    Q  = V ./ (W H)  [+ cost sums]   (152, 210)
//   A = V H'   B = W G_H           (149-150)        [cost of the previous iteration + stop test]
//   a_k = <W_k,A_k>  b_k = <W_k,B_k>                R  = Q H'   c_k = <W_k,R_k>      (152-153)
//   W <- W.*(A + W b)./max(B + W a + lW, eps) (168) W <- W.*(R + W hs ws)./max(hs + W c + lW, eps)
//   W <- W diag(1/|W_k|)               (169)        W <- W diag(1/|W_k|)
//   G_W = W' W                                      Q  = V ./ (W H)                  (173, 183)
//   H <- H.*(W'V)./max(G_W H + lH, eps) (180-199)   H <- H.*(W'Q)./max(ws + lH, eps) (183-199)
//
// The Euclidean cost 0.5*|V - W H|^2 (208) is evaluated through
// 0.5*(|V|^2 - 2<W'V, H> + <G_W, G_H>) with the Gram matrices the updates need
// anyway (or explicitly with cost_mode = NMFB_COST_DIRECT); it becomes
// available once G_H of the *new* H exists, i.e. at the start of the next
// iteration, and is finalised by one trailing Gram product after the last.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "comm.cuh"
#include "engine.cuh"
#include "ew_kernels.cuh"
#include "w_shard.cuh"

using namespace nmfb;

// SMs a contraction with tail helpers leaves to the Gram product running beside it (two CTA pairs)
static constexpr int kTailReserveSms = 4;

struct NmfSession {
  Arena ar;
  int K = 0, Kp = 0, m = 0, n = 0;
  long long ldw = 0, ldh = 0;
  int divergence = NMFB_DIV_EUCLIDEAN;
  bool W_fixed = false, H_fixed = false, direct_cost = false;
  bool overlap = false;  // the Gram products (and the cost) run on the side stream BESIDE the two large
                         // contractions, which wait for them at a device-side gate before their short phase 1
  bool gate_h = false;   // H-step contraction small enough to leave SMs for gram(W)
  bool side_gh = false;  // several GPUs: gram(H) runs beside the A GEMM, joined before the all-reduce
  unsigned int* gates = nullptr;  // [0] G_H ready for iteration i (value i+1), [1] G_W ready
  bool h_split = false;  // too few sample tiles for the fused H update: split-K GEMM + h_finish
  bool tail_a = false, tail_h = false;  // the A / H-step contraction runs with tail helpers on the idle SMs
  int h_tile_n = 0;      // fused H update with narrower tiles (more CTAs) instead of split-K
  float *Nbuf = nullptr, *Dbuf = nullptr;
  float lambda_w = 0.f, lambda_h = 0.f;
  // per-basis settings of a multi-source run (nmfb_config::*_k); null = the scalars apply
  bool per_basis = false;
  float *lamW_k = nullptr, *lamH_k = nullptr;
  int *fixW_k = nullptr, *fixH_k = nullptr;
  int maxiter = 100;
  double tolerance = 1e-3;
  int iters_enqueued = 0;
  bool finalized = false;
  double device_ms = 0.0;

  float *Wm = nullptr, *Wt = nullptr, *Hm = nullptr, *Ht = nullptr;
  float *A = nullptr, *B = nullptr, *Q = nullptr;
  float *pcoef = nullptr, *qcoef = nullptr, *bvec = nullptr, *wsf = nullptr;
  double *ab = nullptr, *norm2 = nullptr, *wsum = nullptr, *hs = nullptr, *scal = nullptr;
  double *cost = nullptr, *vstats = nullptr;
  int* stop = nullptr;
  unsigned int* ticket = nullptr;
  double vsq = 0.0;
  const float* Vmma = nullptr;  // V as the tensor cores read it (tf32-rounded copy)

  GramOp gramH, gramW;
  GemmOp gemmA, gemmB, gemmH, gemmS, gemmR;
  // IS / AB divergences ("two-weight" updates: both gradients are contractions with an element-wise
  // function of V and V_hat, nmf.m:154-164,185-195)
  bool lnmf = false;     // lnmf.m: KL-type updates with unit-sum bases and a square-root H step
  bool two_weight = false;
  float* Q2 = nullptr;   // Qp next to Q = Qn
  GemmOp gemmRb, gemmHn, gemmHd;
  float expo = 0.f;      // outer exponent of the AB gradients (1/alpha, dual: 1/beta)
  double ab_scale = 0.0; // -1/(alpha beta) of the AB cost (nmf.m:214)
  KlOp klW, klH;         // fused KL halves (kl_fused.cuh)
  bool kl_fused = false;
  AbOp abW, abH;         // fused IS / AB halves (ab_fused.cuh)
  bool ab_fused = false;
  bool kl_store_n = false;  // unfused KL whose H step needs N = W'Q as a matrix (lnmf, per-source settings, tied Z)
  // constrainednmf.m: H = Z*A, A the 0/1 label matrix (columns of H tied to columns of Z)
  bool tied = false;
  int nz = 0;
  long long ldz = 0;
  float* Zm = nullptr;
  int *col2z = nullptr, *seg = nullptr;
  // multi-GPU, row-sharded W step (w_shard.cuh): this rank's rows of W, byte offsets inside the peer region
  bool w_sharded = false;
  bool ws_open_barrier = false;  // the small all-reduce runs on the side stream: the sharded kernel meets the peers itself
  unsigned long long* ws_timing = nullptr;  // NMFB_WS_TIMING=1: per-block phase stamps of the last sharded W step
  int ws_grid = 0;
  int r0 = 0, mb = 0;
  size_t a_off = 0, b_off = 0, wt_off = 0, wm_off = 0, x_off = 0;
  float* packed = nullptr;  // multi-GPU: [A | G_H] contiguous fp32 for the single all-reduce
  char* region = nullptr;   // multi-GPU: allocation shared with the peers ([packed | hs | scal | flags])
  int* pinned = nullptr;    // host copy of stop[0..1]
};

namespace nmfdetail {

struct D2FArgs {
  const double* src;
  float* dst;
  int n;
};
__global__ void d2f_kernel(D2FArgs a, const int* stop) {
  NMFB_STOP_GUARD(stop);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n) a.dst[i] = static_cast<float>(a.src[i]);
}

int zero_async(nmfb_handle* h, void* p, size_t bytes) {
  NMFB_CUDA(h, cudaMemsetAsync(p, 0, bytes, h->stream));
  return NMFB_OK;
}

int normalize_defaults(const nmfb_config* in, nmfb_config* out) {
  std::memset(out, 0, sizeof(*out));
  if (in) *out = *in;
  if (out->maxiter <= 0) out->maxiter = 100;           // nmf.m:404-406
  if (!(out->tolerance > 0)) out->tolerance = 1e-3;    // nmf.m:409-411
  if (out->W_sparsity < 0) out->W_sparsity = 0;        // nmf.m:321-333
  if (out->H_sparsity < 0) out->H_sparsity = 0;
  return 0;
}

}  // namespace nmfdetail
using namespace nmfdetail;

// ------------------------------------------------------------------ IS / AB divergences
// Per half iteration: V_hat = W H on the tensor cores with the two weight matrices
//   IS        Qn = V ./ V_hat.^2                  Qp = 1 ./ V_hat                  (nmf.m:155-156)
//   AB        Qn = V.^a .* V_hat.^(b-1)           Qp = V_hat.^(a+b-1)              (nmf.m:162-163)
//   AB, a = 0 Qn = V.^(a-1) .* V_hat.^b           Qp = V.^(a+b-1)                  (nmf.m:159-160)
// written by the epilogue (EPI_ABQ), then  W: A = Qn H', B = Qp H'  and the Euclidean-shaped W step
// (neg = A + W diag(<W_k,B_k>), pos = B + W diag(<W_k,A_k>), both raised to 1/a or 1/b for AB);
// H: N = W' Qn, D = W' Qp, H <- H .* N^e ./ max(D^e + lambda, eps).  The divergence itself
// (nmf.m:211-214) is summed by the same epilogue one V_hat later, as for KL.
static int plan_two_weight(nmfb_handle* h, NmfSession* s, const nmfb_config& cfg) {
  Arena* ar = &s->ar;
  const int m = s->m, n = s->n, Kp = s->Kp;
  const int* stop = s->stop;
  int mode = ABQ_IS;
  if (s->divergence == NMFB_DIV_AB) {
    const bool dual = cfg.alpha == 0;  // nmf.m:124-128
    mode = dual ? ABQ_AB_DUAL : ABQ_AB;
    s->expo = static_cast<float>(1.0 / (dual ? cfg.beta : cfg.alpha));
    s->ab_scale = -1.0 / (cfg.alpha * cfg.beta);
  }
  NMFB_TRY(ar->alloc(h, &s->Nbuf, static_cast<size_t>(Kp) * s->ldh));
  NMFB_TRY(ar->alloc(h, &s->Dbuf, static_cast<size_t>(Kp) * s->ldh));
  {
    // Fused path (ab_fused.cuh): V_hat and the two weight matrices stay on chip.  Needs K <= 128 and a row-major
    // copy of V for the H half (one m x n buffer instead of the two the unfused path keeps for Qn and Qp).
    size_t free_b = 0, total_b = 0;
    const long long ldn = round_up(n, 4);
    const size_t need = static_cast<size_t>(m) * ldn * sizeof(float);
    const char* env = std::getenv("NMFB_AB_UNFUSED");
    if (!(env && env[0] == '1') && Kp <= kKlMaxKp && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
        free_b + h->pool.held > need + (size_t(1) << 30)) {
      float* Vrm = nullptr;
      NMFB_TRY(ar->alloc(h, &Vrm, static_cast<size_t>(m) * ldn));
      dim3 grid((n + 31) / 32, (m + 31) / 32);
      transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(h->Vraw, h->ldv, Vrm, ldn, m, n);
      NMFB_TRY(check_launch(h, "transpose(V)"));
      const float al = static_cast<float>(cfg.alpha), be = static_cast<float>(cfg.beta);
      NMFB_TRY(plan_ab(h, ar, &s->abW, s->Wt, s->ldw, s->Ht, s->ldh, h->Vraw, h->ldv, m, n, Kp, mode, al, be, stop));
      NMFB_TRY(plan_ab(h, ar, &s->abH, s->Ht, s->ldh, s->Wt, s->ldw, Vrm, ldn, n, m, Kp, mode, al, be, stop));
      s->abW.args.scal = s->scal + 2;
      s->ab_fused = true;
      return NMFB_OK;
    }
  }
  NMFB_TRY(ar->alloc(h, &s->Q, static_cast<size_t>(n) * h->ldv));
  NMFB_TRY(ar->alloc(h, &s->Q2, static_cast<size_t>(n) * h->ldv));
  MatRef Xs{s->Wt, m, Kp, s->ldw, true};
  MatRef Ys{s->Ht, n, Kp, s->ldh, true};
  NMFB_TRY(plan_fused(h, &s->gemmS, EPI_ABQ, Xs, Ys, Kp, nullptr, nullptr, 0, m, round_up(n, 64), n, stop));
  GemmArgs& q = s->gemmS.L.args;
  q.Vsrc = h->Vraw;
  q.Qout = s->Q;
  q.Qout2 = s->Q2;
  q.ldv = h->ldv;
  q.scal = s->scal + 2;
  q.ab_mode = mode;
  q.ab_alpha = static_cast<float>(cfg.alpha);
  q.ab_beta = static_cast<float>(cfg.beta);
  {
    std::string pe = set_v_prefetch(&s->gemmS.L, h->Vraw, m, n, h->ldv);
    if (!pe.empty()) return h->fail(NMFB_ERR_CUDA, "%s", pe.c_str());
  }
  // W step: A = Qn H', B = Qp H'
  MatRef Yh{s->Ht, n, Kp, s->ldh, false};
  const int tilesW = (m + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
  const bool splitW = tilesW * 2 <= h->num_sms;
  MatRef Xn{s->Q, m, n, h->ldv, true};
  MatRef Xp{s->Q2, m, n, h->ldv, true};
  NMFB_TRY(plan_store(h, ar, &s->gemmR, Xn, Yh, n, nullptr, nullptr, 0, m, Kp, s->A, nullptr, s->ldw, splitW, stop));
  NMFB_TRY(plan_store(h, ar, &s->gemmRb, Xp, Yh, n, nullptr, nullptr, 0, m, Kp, s->B, nullptr, s->ldw, splitW, stop));
  // H step: N = W' Qn, D = W' Qp
  MatRef Yw{s->Wt, m, Kp, s->ldw, false};
  const int tilesH = (n + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
  const bool splitH = tilesH * 2 <= h->num_sms;
  MatRef Xnt{s->Q, m, n, h->ldv, false};
  MatRef Xpt{s->Q2, m, n, h->ldv, false};
  NMFB_TRY(plan_store(h, ar, &s->gemmHn, Xnt, Yw, m, nullptr, nullptr, 0, n, Kp, s->Nbuf, nullptr, s->ldh, splitH, stop));
  NMFB_TRY(plan_store(h, ar, &s->gemmHd, Xpt, Yw, m, nullptr, nullptr, 0, n, Kp, s->Dbuf, nullptr, s->ldh, splitH, stop));
  return NMFB_OK;
}

// ------------------------------------------------------------------ setup
static int nmf_setup(nmfb_handle* h, NmfSession* s, int K, const nmfb_config* cfg_in, bool lnmf = false) {
  if (h->Vraw == nullptr) return h->fail(NMFB_ERR_NO_DATA, "nmf: call nmfb_set_V first");
  if (K <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "nmf: num_basis_elems must be positive");
  nmfb_config cfg;
  normalize_defaults(cfg_in, &cfg);
  s->lnmf = lnmf;
  if (lnmf) {  // lnmf.m has no divergence / sparsity options (lnmf.m:95-135)
    cfg.divergence = NMFB_DIV_KL;
    cfg.W_sparsity = cfg.H_sparsity = 0;
    cfg.W_sparsity_k = cfg.H_sparsity_k = nullptr;
    cfg.W_fixed_k = cfg.H_fixed_k = nullptr;
    if (comm_size(h->comm) > 1) return h->fail(NMFB_ERR_UNSUPPORTED, "lnmf: one GPU only");
  }
  switch (cfg.divergence) {
    case NMFB_DIV_EUCLIDEAN:
    case NMFB_DIV_KL:
      break;
    case NMFB_DIV_AB:
      if (cfg.alpha == 0 && cfg.beta == 0)  // nmf.m:120-122
        return h->fail(NMFB_ERR_AB_ZERO, "alpha = 0 and beta = 0 is not supported at this time.");
      [[fallthrough]];
    case NMFB_DIV_IS:
      break;
    default:  // nmf.m:165-166 ('frobenius' included: nmf.m has no such case)
      return h->fail(NMFB_ERR_DIVERGENCE,
                     "No update equations defined for cost function with divergence type %d",
                     cfg.divergence);
  }
  const int m = h->m, n = h->n;
  const bool trace = std::getenv("NMFB_TRACE") != nullptr;
  timespec tr0;
  clock_gettime(CLOCK_MONOTONIC, &tr0);
  auto lap = [&](const char* what) {  // NMFB_TRACE=1: where the setup time of a call goes (synchronises)
    if (!trace) return;
    cudaStreamSynchronize(h->stream);
    timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "[nmfb] nmf setup: %s %.1f ms\n", what, (t1.tv_sec - tr0.tv_sec) * 1e3 + (t1.tv_nsec - tr0.tv_nsec) * 1e-6);
    tr0 = t1;
  };
  s->K = K;
  s->Kp = round_up(K, 32);
  s->m = m;
  s->n = n;
  s->ldw = round_up(m, 4);
  s->ldh = round_up(n, 4);
  s->divergence = cfg.divergence;
  s->W_fixed = cfg.W_fixed != 0;
  s->H_fixed = cfg.H_fixed != 0;
  s->direct_cost = cfg.cost_mode == NMFB_COST_DIRECT;
  s->lambda_w = static_cast<float>(cfg.W_sparsity);
  s->lambda_h = static_cast<float>(cfg.H_sparsity);
  s->maxiter = cfg.maxiter;
  s->tolerance = cfg.tolerance;
  const int Kp = s->Kp;
  // ---- per-basis overrides (multi-source runs): collapse to the scalars when they are uniform
  std::vector<float> lw(Kp, 0.f), lh(Kp, 0.f);
  std::vector<int> fw(Kp, 1), fh(Kp, 1);  // padding rows/columns are never touched
  if (cfg.W_sparsity_k || cfg.H_sparsity_k || cfg.W_fixed_k || cfg.H_fixed_k) {
    bool uniform = true;
    for (int k = 0; k < K; ++k) {
      lw[k] = static_cast<float>(cfg.W_sparsity_k ? std::max(0.0, cfg.W_sparsity_k[k]) : cfg.W_sparsity);
      lh[k] = static_cast<float>(cfg.H_sparsity_k ? std::max(0.0, cfg.H_sparsity_k[k]) : cfg.H_sparsity);
      fw[k] = cfg.W_fixed_k ? (cfg.W_fixed_k[k] != 0) : (cfg.W_fixed != 0);
      fh[k] = cfg.H_fixed_k ? (cfg.H_fixed_k[k] != 0) : (cfg.H_fixed != 0);
      uniform = uniform && lw[k] == lw[0] && lh[k] == lh[0] && fw[k] == fw[0] && fh[k] == fh[0];
    }
    if (uniform) {
      s->lambda_w = lw[0];
      s->lambda_h = lh[0];
      s->W_fixed = fw[0] != 0;
      s->H_fixed = fh[0] != 0;
    } else {
      s->per_basis = true;
      s->W_fixed = s->H_fixed = false;  // the masks decide per basis; every kernel of the iteration runs
    }
  }
  const bool kl = s->divergence == NMFB_DIV_KL;
  const bool tw = s->divergence == NMFB_DIV_IS || s->divergence == NMFB_DIV_AB;
  s->two_weight = tw;
  Arena* ar = &s->ar;
  std::vector<int> seg_host;
  if (h->tie_col2z != nullptr) {  // nmfb_constrainednmf: columns of H are tied through the label matrix
    s->tied = true;
    s->nz = h->tie_nz;
    s->ldz = round_up(s->nz, 4);
    if (comm_size(h->comm) > 1) return h->fail(NMFB_ERR_UNSUPPORTED, "constrainednmf: one GPU only");
    if (s->per_basis) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "constrainednmf has no per-source settings");
    if (s->nz <= 0 || s->nz > n) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "constrainednmf: bad number of Z columns");
    seg_host.assign(static_cast<size_t>(s->nz) + 1, 0);
    for (int j = 0; j < n; ++j) {
      const int z = h->tie_col2z[j];
      if (z < 0 || z >= s->nz || (j > 0 && z < h->tie_col2z[j - 1]))
        return h->fail(NMFB_ERR_INVALID_ARGUMENT, "constrainednmf: the column map must be non-decreasing in [0, nz)");
      seg_host[z + 1] = j + 1;
    }
    for (int z = 0; z < s->nz; ++z) {
      if (seg_host[z + 1] == 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "constrainednmf: column %d of Z has no sample", z);
    }
  }

  NMFB_TRY(ar->alloc(h, &s->Hm, static_cast<size_t>(Kp) * s->ldh));
  NMFB_TRY(ar->alloc(h, &s->Ht, static_cast<size_t>(Kp) * s->ldh));
  // packed = [A (Kp x ldw) | G_H (Kp x Kp, KL: unused) | ...]: one buffer so that a
  // multi-GPU run can all-reduce it in a single call.
  // With several GPUs the doubles that travel with it (hs, scal) and the barrier flags of the
  // peer-memory all-reduce live behind it in the SAME allocation, owned by the communicator and
  // mapped into the other ranks (comm_acquire_region).
  // (IS / AB: [A | B], both W-step matrices are sums over the column shards)
  const size_t packed_floats =
      static_cast<size_t>(Kp) * s->ldw + (tw ? static_cast<size_t>(Kp) * s->ldw : static_cast<size_t>(Kp) * Kp);
  const bool share = comm_size(h->comm) > 1;
  const size_t dbl_off = (packed_floats * sizeof(float) + 255) / 256 * 256;
  if (share) {
    // region = [packed | hs, scal | W tf32 | W fp32 | exchange slots of the sharded W step]
    const size_t w_bytes = static_cast<size_t>(Kp) * s->ldw * sizeof(float);
    const size_t wt_off = (dbl_off + (static_cast<size_t>(Kp) + 8) * sizeof(double) + 255) / 256 * 256;
    const size_t wm_off = wt_off + (w_bytes + 255) / 256 * 256;
    const size_t x_off = wm_off + (w_bytes + 255) / 256 * 256;
    const size_t total = x_off + 2 * static_cast<size_t>(kMaxBlocks) * kMaxRanks * 32;
    char* region = nullptr;
    NMFB_TRY(comm_acquire_region(h, total, &region));
    s->packed = reinterpret_cast<float*>(region);
    s->hs = reinterpret_cast<double*>(region + dbl_off);
    s->scal = s->hs + Kp;
    s->region = region;
    s->Wt = reinterpret_cast<float*>(region + wt_off);
    s->Wm = reinterpret_cast<float*>(region + wm_off);
    // Row-sharded W step over peer memory: rank r owns rows [r0, r0 + mb) of W (in units of 4 rows,
    // padding rows of the leading dimension included).  Needs the peer mapping (else NCCL + the
    // replicated W step) and a row block that fits the kernel's registers.
    const int nr = comm_size(h->comm), rk = comm_rank(h->comm);
    const int rb = round_up((static_cast<int>(s->ldw) + nr - 1) / nr, 4);
    s->r0 = std::min(static_cast<int>(s->ldw), rk * rb);
    s->mb = std::min(static_cast<int>(s->ldw), s->r0 + rb) - s->r0;
    const char* env = std::getenv("NMFB_W_SHARD");
    s->w_sharded = comm_peer_table(h, nullptr) && !lnmf && !(env && env[0] == '0') && rb <= 4 * kWsCache * kWsThreads;
    if (s->w_sharded && std::getenv("NMFB_WS_TIMING")) NMFB_TRY(ar->alloc(h, &s->ws_timing, 8 * kMaxBlocks));
    const size_t base = comm_region_offset(h, region);
    s->a_off = base;
    s->b_off = tw ? base + w_bytes : 0;
    s->wt_off = base + wt_off;
    s->wm_off = base + wm_off;
    s->x_off = base + x_off;
  } else {
    NMFB_TRY(ar->alloc(h, &s->packed, packed_floats));
    NMFB_TRY(ar->alloc(h, &s->Wm, static_cast<size_t>(Kp) * s->ldw));
    NMFB_TRY(ar->alloc(h, &s->Wt, static_cast<size_t>(Kp) * s->ldw));
  }
  s->A = s->packed;
  if (tw) s->B = s->packed + static_cast<size_t>(Kp) * s->ldw;
  else NMFB_TRY(ar->alloc(h, &s->B, static_cast<size_t>(Kp) * s->ldw));
  NMFB_TRY(ar->alloc(h, &s->pcoef, Kp));
  NMFB_TRY(ar->alloc(h, &s->qcoef, Kp));
  NMFB_TRY(ar->alloc(h, &s->bvec, Kp));
  NMFB_TRY(ar->alloc(h, &s->wsf, Kp));
  NMFB_TRY(ar->alloc(h, &s->ab, 4 * Kp));  // [ab (2 Kp) | norm2 (Kp) | wsum (Kp)]: WStepArgs::acc
  s->norm2 = s->ab + 2 * Kp;
  s->wsum = s->ab + 3 * Kp;
  NMFB_TRY(ar->alloc(h, &s->ticket, 2));
  if (s->per_basis) {
    NMFB_TRY(ar->alloc(h, &s->lamW_k, Kp));
    NMFB_TRY(ar->alloc(h, &s->lamH_k, Kp));
    NMFB_TRY(ar->alloc(h, &s->fixW_k, Kp));
    NMFB_TRY(ar->alloc(h, &s->fixH_k, Kp));
    NMFB_CUDA(h, cudaMemcpyAsync(s->lamW_k, lw.data(), Kp * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    NMFB_CUDA(h, cudaMemcpyAsync(s->lamH_k, lh.data(), Kp * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    NMFB_CUDA(h, cudaMemcpyAsync(s->fixW_k, fw.data(), Kp * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NMFB_CUDA(h, cudaMemcpyAsync(s->fixH_k, fh.data(), Kp * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
  }
  if (!share) {
    NMFB_TRY(ar->alloc(h, &s->hs, Kp));
    NMFB_TRY(ar->alloc(h, &s->scal, 8));
  }
  NMFB_TRY(ar->alloc(h, &s->cost, static_cast<size_t>(s->maxiter) + 1));
  NMFB_TRY(ar->alloc(h, &s->stop, 2));
  s->pinned = h->pinned;
  s->pinned[0] = s->pinned[1] = 0;

  lap("device blocks for the factors");
  // ---- initial factors (nmf.m:130-134; defaults nmf.m:277,298)
  {
    std::vector<float> tmp;
    const float* Wsrc = cfg.W_init;
    if (!Wsrc) {
      tmp.resize(static_cast<size_t>(m) * K);
      fill_uniform(tmp, cfg.seed * 2 + 1, true);
      Wsrc = tmp.data();
    }
    NMFB_TRY(upload_colmajor(h, Wsrc, m, K, s->Wm, s->ldw));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    const float* Hsrc = cfg.H_init;
    if (!Hsrc) {
      tmp.resize(static_cast<size_t>(K) * n);
      fill_uniform(tmp, cfg.seed * 2 + 2, true);
      Hsrc = tmp.data();
    }
    if (s->tied) {
      // constrainednmf.m:174-177: Z (given, or rand) and H = Z*A
      NMFB_TRY(ar->alloc(h, &s->Zm, static_cast<size_t>(Kp) * s->ldz));
      NMFB_TRY(ar->alloc(h, &s->col2z, n));
      NMFB_TRY(ar->alloc(h, &s->seg, static_cast<size_t>(s->nz) + 1));
      NMFB_CUDA(h, cudaMemcpyAsync(s->col2z, h->tie_col2z, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      NMFB_CUDA(h, cudaMemcpyAsync(s->seg, seg_host.data(), seg_host.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      const float* Zsrc = h->tie_Zinit;
      if (!Zsrc) {
        tmp.resize(static_cast<size_t>(K) * s->nz);
        fill_uniform(tmp, cfg.seed * 2 + 2, false);
        Zsrc = tmp.data();
      }
      NMFB_TRY(upload_H(h, ar, Zsrc, K, s->nz, s->Zm, s->ldz));
      tied_gather_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Zm, s->ldz, s->col2z, s->Hm, s->Ht, s->ldh, n, nullptr);
      NMFB_TRY(check_launch(h, "tied_gather(H init)"));
    } else {
      NMFB_TRY(upload_H(h, ar, Hsrc, K, n, s->Hm, s->ldh));
    }
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  // W columns -> unit L2 (also for a user-supplied W_init, nmf.m:133); H is not rescaled
  // (lnmf.m:63: unit column SUM instead)
  vec_sums_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm, K, m, s->ldw, lnmf ? s->norm2 : nullptr,
                                                         lnmf ? nullptr : s->norm2, nullptr);
  NMFB_TRY(check_launch(h, "vec_sums(W init)"));
  w_normalize_kernel<<<vec_grid(m, K), 256, 0, h->stream>>>(s->Wm, s->Wt, m, s->ldw, K, 1, lnmf ? 2 : 0, s->norm2,
                                                            s->wsum, nullptr, nullptr);
  NMFB_TRY(check_launch(h, "w_normalize(init)"));
  round_copy_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, s->Ht, K, n, s->ldh, nullptr);
  NMFB_TRY(check_launch(h, "round_copy(H init)"));
  NMFB_TRY(zero_async(h, s->scal, 8 * sizeof(double)));
  vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, s->ldh, s->hs, nullptr, nullptr);
  NMFB_TRY(check_launch(h, "vec_sums(H init)"));

  lap("upload + normalise W_init, H_init");
  // ---- V
  if (tw) {
    // the weights are formed from the fp32 V in the epilogue of V_hat = W H; nothing to prepare
  } else if (kl) {
    VStats st;
    NMFB_TRY(compute_v_stats(h, true, &st, &s->vstats, nullptr, ar));
    NMFB_TRY(comm_allreduce(h, nullptr, 0, s->vstats, 4, nullptr, 0));  // global sums over all shards
  } else {
    double* sq = nullptr;
    NMFB_TRY(ar->alloc(h, &sq, 1));
    NMFB_TRY(prepare_v_work(h, false, true, sq, nullptr));
    NMFB_TRY(comm_allreduce(h, nullptr, 0, sq, 1, nullptr, 0));  // |V|^2 over all column shards
    NMFB_CUDA(h, cudaMemcpyAsync(&s->vsq, sq, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    s->Vmma = h->Vwork;
  }

  lap("V statistics / tf32 working copy");
  // ---- plan the contractions
  const int* stop = s->stop;
  const bool multi = comm_size(h->comm) > 1;
  if (tw) return plan_two_weight(h, s, cfg);
  if (!kl) {
    // decide up front which Gram products run beside a large contraction (see below)
    const int tilesA = (m + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    const int ctasA = ((m + 2 * kTileM - 1) / (2 * kTileM)) * 2 * ((Kp + kMaxN - 1) / kMaxN);
    const int ctasH = ((n + 2 * kTileM - 1) / (2 * kTileM)) * 2 * ((Kp + kMaxN - 1) / kMaxN);
    const char* env = std::getenv("NMFB_OVERLAP");
    auto helpers_for = [&](int epi, int rows, long long kdim) {
      if (multi || Kp > kMaxN) return 0;
      GemmLaunch d;
      std::memset(&d, 0, sizeof(d));
      d.cg = choose_cg(epi, rows, Kp, false, false, 0);
      d.grid = dim3(static_cast<unsigned>((rows + 2 * kTileM - 1) / (2 * kTileM)) * 2, 1, 1);
      d.args.nkb0 = d.args.nkb_seg = static_cast<int>((kdim + kBlockK - 1) / kBlockK);
      int kp = 0;
      return plan_tail_helpers(d, epi, h->num_sms, kTailReserveSms, &kp);
    };
    s->h_split = ctasH * 2 <= h->num_sms && n > kTileM;
    // Experiment (NMFB_H_TAIL=1): 25 - 37 sample tiles (a 2-GPU shard of the north-star problem has 32) with one
    // helper pair per tile instead of split-K + h_finish: the fused H update is kept, N and D never exist in HBM.
    // Measured at 16384 x 8192, K = 256: 153.6 us against 149.5 us for split-K + h_finish - no gain, off by default.
    if (const char* e5 = std::getenv("NMFB_H_TAIL"))
      if (e5[0] == '1' && s->h_split && helpers_for(EPI_HUPDATE, n, m) > 0) s->h_split = false;
    // Experiment (NMFB_H_TILEN=<width>): with few sample tiles, narrower tiles instead of split-K - twice the
    // CTAs, each still runs the whole contraction and keeps the fused H update.  Measured at 2 GPUs
    // (16384 x 8192 shard, K = 256): 222 us against 144 us for split-K + h_finish - every CTA streams the V
    // tile for half the tensor work, so the kernel becomes load-bound.  Off by default.
    int& h_tile_n = s->h_tile_n;
    h_tile_n = 0;
    if (const char* e3 = std::getenv("NMFB_H_TILEN")) h_tile_n = std::atoi(e3);
    if (h_tile_n > 0) s->h_split = false;
    if (const char* e2 = std::getenv("NMFB_H_SPLIT")) s->h_split = e2[0] == '1';  // tests force either path
    if (s->per_basis) s->h_split = true;  // per-basis lambda / fixed rows live in h_finish, not in the fused epilogue
    if (s->tied) s->h_split = true;       // the Z step sums N and D over the samples of a class (tied_update_kernel)
    s->overlap = !multi && !s->direct_cost && !s->W_fixed && !s->H_fixed && !(tilesA * 2 <= h->num_sms) &&
                 !(env && env[0] == '0') && m > kTileM && ctasA + 8 <= h->num_sms;
    const bool can_side = !s->direct_cost && !s->W_fixed && !s->H_fixed && !(env && env[0] == '0') && m > kTileM &&
                          ctasA + 8 <= h->num_sms;
    // The H-step contraction may only wait at the gate for gram(W) when its whole grid is resident with
    // SMs to spare: gram(W) (<= 20 CTAs with large dynamic shared memory) needs free SMs to run on, and
    // a grid of more CTAs than SMs spinning at the gate would never let it start (deadlock).  The fused
    // kernel launches ctasH CTAs; the split-K variant is planned below with num_sms - 20 as its budget,
    // which it can only honour when the unsplit grid already fits (choose_splits never goes below 1).
    const int ctasHf = h_tile_n > 0 ? ctasH * ((Kp + h_tile_n - 1) / h_tile_n) : ctasH;  // fused kernel's grid
    const bool h_grid_fits = s->h_split ? ctasH + 20 <= h->num_sms : ctasHf + 8 <= h->num_sms;
    s->gate_h = ((s->overlap && !s->h_split) || (can_side && s->h_split)) && h_grid_fits;
    // Tail helpers (panel_gemm.cuh, GemmArgs::sk_*): at the north-star shape each of the two contractions has 64
    // pair tiles - 128 of 148 SMs.  Helper pairs take the tails of the contractions so that 144 SMs work through
    // the whole launch; the Gram product running beside it is then planned for the 4 SMs that stay free.
    s->tail_a = s->overlap && helpers_for(EPI_STORE, m, n) > 0;
    s->tail_h = s->gate_h && !s->h_split && h_tile_n == 0 && helpers_for(EPI_HUPDATE, n, m) > 0;
    s->side_gh = multi && can_side;
    {
      const char* e4 = std::getenv("NMFB_WS_SIDE");
      s->ws_open_barrier = s->w_sharded && s->side_gh && !s->direct_cost && !(e4 && e4[0] == '0');
    }
  }
  NMFB_TRY(plan_gram(h, ar, &s->gramW, s->Wt, Kp, m, s->ldw, stop, nullptr, 0,
                     s->gate_h ? (s->tail_h ? kTailReserveSms : 20) : 0));
  if (!kl) {
    NMFB_TRY(plan_gram(h, ar, &s->gramH, s->Ht, Kp, n, s->ldh, stop, nullptr, 0,
                       (s->overlap || s->side_gh) ? (s->tail_a ? kTailReserveSms : 20) : 0));
    if (multi) {  // G_H must sit behind A in the packed buffer
      s->gramH.g32 = s->packed + static_cast<size_t>(Kp) * s->ldw;
    }
    // A = V H'  (X = V, rows i contiguous -> MN-major; Y = H rows, K-major), B = W G_H
    MatRef Xv{s->Vmma, m, n, h->ldv, true};
    MatRef Yh{s->Ht, n, Kp, s->ldh, false};
    MatRef Xw{s->Wt, m, Kp, s->ldw, true};
    MatRef Yg{s->gramH.gtf, Kp, Kp, Kp, false};
    const int tiles = (m + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    const bool split = tiles * 2 <= h->num_sms;
    {
      // B = W G_H is the short phase 1 of the A GEMM, so G_H is not needed until the long phase 0
      // is over: gram(H), <G_W,G_H>, the cost and the stop test run on the side stream on SMs the
      // GEMM leaves idle, and the GEMM's producer waits at a gate before its first phase-1 load.
      // (A grid that fills every SM would leave the side stream nothing to run on: no overlap then.)
      NMFB_TRY(ar->alloc(h, &s->gates, 2));
    }
    if (multi || split) {
      NMFB_TRY(plan_store(h, ar, &s->gemmA, Xv, Yh, n, nullptr, nullptr, 0, m, Kp, s->A, nullptr,
                          s->ldw, split, stop));
      if (s->w_sharded) {  // B = W G_H only on the rows whose W step this rank takes
        const int rows = std::max(0, std::min(m, s->r0 + s->mb) - s->r0);
        if (rows > 0) {
          MatRef Xwb{s->Wt + s->r0, rows, Kp, s->ldw, true};
          NMFB_TRY(plan_store(h, ar, &s->gemmB, Xwb, Yg, Kp, nullptr, nullptr, 0, rows, Kp, s->B + s->r0, nullptr,
                              s->ldw, false, stop));
        }
      } else {
        NMFB_TRY(plan_store(h, ar, &s->gemmB, Xw, Yg, Kp, nullptr, nullptr, 0, m, Kp, s->B, nullptr,
                            s->ldw, false, stop));
      }
    } else {
      NMFB_TRY(plan_store(h, ar, &s->gemmA, Xv, Yh, n, &Xw, &Yg, Kp, m, Kp, s->A, s->B, s->ldw, false,
                          stop));
      if (s->overlap) s->gemmA.L.args.gate = s->gates + 0;
      if (s->tail_a) NMFB_TRY(enable_tail_helpers(h, ar, &s->gemmA, kTailReserveSms));
    }
    // H update: N = W'V, D = G_W H (X = H, columns j contiguous -> MN-major).
    // Reading V with the contraction index contiguous (K-major) makes every TMA row a lone
    // 128-byte DRAM access 64 KB away from the next one; with a row-major copy of V the same
    // product reads 512-byte runs (MN-major X), worth ~15 % of this kernel.  The copy is made
    // once per call when memory allows (it costs one more m x n buffer).
    MatRef Xvt{s->Vmma, m, n, h->ldv, false};
    {
      size_t free_b = 0, total_b = 0;
      const long long ldn = round_up(n, 4);
      const size_t need = static_cast<size_t>(m) * ldn * sizeof(float);
      const char* env = std::getenv("NMFB_NO_VT");
      if (!(env && env[0] == '1') && m >= 1024 && n >= 1024 &&
          cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b + h->pool.held > need + (size_t(2) << 30)) {
        float* Vrm = nullptr;
        NMFB_TRY(ar->alloc(h, &Vrm, static_cast<size_t>(m) * ldn));
        dim3 grid((n + 31) / 32, (m + 31) / 32);
        // Vrm[i][j] = V[j][i]  (src is [n][ldv])
        transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(s->Vmma, h->ldv, Vrm, ldn, m, n);
        NMFB_TRY(check_launch(h, "transpose(V)"));
        Xvt = MatRef{Vrm, n, m, ldn, true};
      }
    }
    MatRef Yw{s->Wt, m, Kp, s->ldw, false};
    MatRef Xh{s->Ht, n, Kp, s->ldh, true};
    MatRef Ygw{s->gramW.gtf, Kp, Kp, Kp, false};
    {
      // The fused H update needs the whole contraction in one CTA pair; with few sample tiles
      // (small column shards on many GPUs) that leaves most SMs idle, so the contraction is
      // split over CTAs instead and the update runs as a separate element-wise kernel.
    }
    if (s->h_split) {
      NMFB_TRY(ar->alloc(h, &s->Nbuf, static_cast<size_t>(Kp) * s->ldh));
      NMFB_TRY(ar->alloc(h, &s->Dbuf, static_cast<size_t>(Kp) * s->ldh));
      const int sms = h->num_sms;
      if (s->gate_h) h->num_sms = std::max(16, sms - 20);  // leave SMs for the Gram product running beside it
      int rc = plan_store(h, ar, &s->gemmH, Xvt, Yw, m, &Xh, &Ygw, Kp, n, Kp, s->Nbuf, s->Dbuf, s->ldh, true, stop);
      h->num_sms = sms;
      NMFB_TRY(rc);
    } else {
    NMFB_TRY(plan_fused(h, &s->gemmH, EPI_HUPDATE, Xvt, Yw, m, &Xh, &Ygw, Kp, n, Kp, Kp, stop, nullptr, s->h_tile_n));
    if (s->tail_h) NMFB_TRY(enable_tail_helpers(h, ar, &s->gemmH, kTailReserveSms));
    }
    if (s->gate_h) {
      const dim3 g = s->gemmH.L.grid;
      const int spare = s->gemmH.L.args.sk_helpers > 0 ? kTailReserveSms : 8;
      if (static_cast<int>(g.x * g.y * g.z) + spare > h->num_sms)  // cannot happen with the guard above
        return h->fail(NMFB_ERR_CUDA, "internal: gated H-step grid of %u CTAs leaves no SM for gram(W)", g.x * g.y * g.z);
      s->gemmH.L.args.gate = s->gates + 1;
    }
    GemmArgs& a = s->gemmH.L.args;
    if (!s->h_split) {
      a.Hm = s->Hm;
      a.Hr32 = s->Ht;
      a.Hc32 = nullptr;
      a.ldh = s->ldh;
      a.lambda = s->lambda_h;
      a.scal = s->scal;
      a.freeze = s->H_fixed ? 1 : 0;
      if (!std::getenv("NMFB_NO_HPREFETCH")) {
        std::string e = set_h_prefetch(&s->gemmH.L, s->Hm, n, Kp, s->ldh);
        if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "%s", e.c_str());
      }
    }
    if (s->direct_cost) {
      MatRef Xs{s->Wt, m, Kp, s->ldw, true};
      MatRef Ys{s->Ht, n, Kp, s->ldh, true};
      NMFB_TRY(plan_fused(h, &s->gemmS, EPI_RESID, Xs, Ys, Kp, nullptr, nullptr, 0, m, round_up(n, 64),
                          n, stop));
      GemmArgs& r = s->gemmS.L.args;
      r.Vsrc = s->Vmma;
      r.ldv = h->ldv;
      r.scal = s->scal + 2;
      std::string pe = set_v_prefetch(&s->gemmS.L, s->Vmma, m, n, h->ldv);
      if (!pe.empty()) return h->fail(NMFB_ERR_CUDA, "%s", pe.c_str());
    }
  } else {
    {
      // Fused path: V_hat and Q = V ./ V_hat stay on chip (kl_fused.cuh).  Needs K <= 128 and a
      // row-major copy of V for the H half.
      size_t free_b = 0, total_b = 0;
      const long long ldn = round_up(n, 4);
      const size_t need = static_cast<size_t>(m) * ldn * sizeof(float);
      const char* env = std::getenv("NMFB_KL_UNFUSED");
      if (!(env && env[0] == '1') && Kp <= kKlMaxKp && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
          free_b + h->pool.held > need + (size_t(1) << 30)) {
        float* Vrm = nullptr;
        NMFB_TRY(ar->alloc(h, &Vrm, static_cast<size_t>(m) * ldn));
        dim3 grid((n + 31) / 32, (m + 31) / 32);
        transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(h->Vraw, h->ldv, Vrm, ldn, m, n);
        NMFB_TRY(check_launch(h, "transpose(V)"));
        NMFB_TRY(plan_kl(h, ar, &s->klW, s->Wt, s->ldw, s->Ht, s->ldh, h->Vraw, h->ldv, m, n, Kp, stop));
        NMFB_TRY(plan_kl(h, ar, &s->klH, s->Ht, s->ldh, s->Wt, s->ldw, Vrm, ldn, n, m, Kp, stop));
        s->klW.args.scal = s->scal + 2;
        s->kl_fused = true;
      }
    }
    // lnmf's square-root step, per-source settings and the label-tied Z step are not forms of the fused
    // H-update epilogue: without the fused KL kernel (K > 128) they take N = W'Q as a stored matrix and
    // finish in kl_h_finish / tied_update
    s->kl_store_n = !s->kl_fused && (s->lnmf || s->tied || s->per_basis);
    if (!s->kl_fused) {
    // Unfused fallback (K > 128 or not enough memory for the row-major copy of V):
    // S = W H (both operands MN-major), Q = V ./ S materialised in HBM
    NMFB_TRY(ar->alloc(h, &s->Q, static_cast<size_t>(n) * h->ldv));
    MatRef Xs{s->Wt, m, Kp, s->ldw, true};
    MatRef Ys{s->Ht, n, Kp, s->ldh, true};
    NMFB_TRY(plan_fused(h, &s->gemmS, EPI_KLQ, Xs, Ys, Kp, nullptr, nullptr, 0, m, round_up(n, 64), n,
                        stop));
    GemmArgs& q = s->gemmS.L.args;
    q.Vsrc = h->Vraw;
    q.Qout = s->Q;
    q.ldv = h->ldv;
    q.scal = s->scal + 2;
    {
      std::string pe = set_v_prefetch(&s->gemmS.L, h->Vraw, m, n, h->ldv);
      if (!pe.empty()) return h->fail(NMFB_ERR_CUDA, "%s", pe.c_str());
    }
    // R = Q H'
    MatRef Xq{s->Q, m, n, h->ldv, true};
    MatRef Yh{s->Ht, n, Kp, s->ldh, false};
    const int tiles = (m + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
    NMFB_TRY(plan_store(h, ar, &s->gemmR, Xq, Yh, n, nullptr, nullptr, 0, m, Kp, s->A, nullptr, s->ldw,
                        tiles * 2 <= h->num_sms, stop));
    // H update: N = W'Q, D = ws
    MatRef Xqt{s->Q, m, n, h->ldv, false};
    MatRef Yw{s->Wt, m, Kp, s->ldw, false};
    if (s->kl_store_n) {
      NMFB_TRY(ar->alloc(h, &s->Nbuf, static_cast<size_t>(Kp) * s->ldh));
      const int tiles_h = (n + kTileM - 1) / kTileM * ((Kp + kMaxN - 1) / kMaxN);
      NMFB_TRY(plan_store(h, ar, &s->gemmH, Xqt, Yw, m, nullptr, nullptr, 0, n, Kp, s->Nbuf, nullptr, s->ldh,
                          tiles_h * 2 <= h->num_sms, stop));
    } else {
    NMFB_TRY(plan_fused(h, &s->gemmH, EPI_HUPDATE, Xqt, Yw, m, nullptr, nullptr, 0, n, Kp, Kp, stop));
    GemmArgs& a = s->gemmH.L.args;
    a.Hm = s->Hm;
    a.Hr32 = s->Ht;
    a.Hc32 = nullptr;
    a.ldh = s->ldh;
    a.lambda = s->lambda_h;
    a.scal = s->scal;
    a.dvec = s->wsf;
    a.freeze = s->H_fixed ? 1 : 0;
    if (!std::getenv("NMFB_NO_HPREFETCH")) {
      std::string e = set_h_prefetch(&s->gemmH.L, s->Hm, n, Kp, s->ldh);
      if (!e.empty()) return h->fail(NMFB_ERR_CUDA, "%s", e.c_str());
    }
    }
    }
  }
  if (trace && !kl)
    fprintf(stderr, "[nmfb] nmf setup: tail helpers A %d pairs (primaries stop at k-block %d of %d), H step %d pairs (%d of %d)\n",
            s->gemmA.L.args.sk_helpers, s->gemmA.L.args.sk_kp, s->gemmA.L.args.nkb0, s->gemmH.L.args.sk_helpers,
            s->gemmH.L.args.sk_kp, s->gemmH.L.args.nkb0);
  if (s->W_fixed) NMFB_TRY(run_gram(h, s->gramW, nullptr));
  D2FArgs da{s->wsum, s->wsf, Kp};
  d2f_kernel<<<(Kp + 127) / 128, 128, 0, h->stream>>>(da, nullptr);
  NMFB_TRY(check_launch(h, "d2f(ws)"));
  lap("plans, tensor maps, row-major copy of V");
  return NMFB_OK;
}

// ------------------------------------------------------------------ one iteration
static int enqueue_cost(nmfb_handle* h, NmfSession* s, int iter, int mode) {
  if (mode == 0) {  // <G_W, G_H> of the trace identity (scal[4] is zero: reset by the previous cost kernel)
    const int cnt = s->Kp * s->Kp;
    gram_dot_kernel<<<std::min(64, (cnt + 1023) / 1024), 256, 0, h->stream>>>(s->gramW.g32, s->gramH.g32, cnt,
                                                                               s->scal + 4, s->stop);
    NMFB_TRY(check_launch(h, "gram_dot"));
  }
  CostArgs c{};
  c.mode = mode;
  c.iter = iter;
  c.Kp = s->Kp;
  c.GW = s->gramW.g32;
  c.GH = s->gramH.g32;
  c.vsq = s->vsq;
  c.vstats = s->vstats;
  c.scal = s->scal;
  c.wsum = s->wsum;
  c.n_wsum = s->Kp;
  c.lambda_w = s->per_basis ? 1.0 : s->lambda_w;
  c.lamw_k = s->lamW_k;
  c.lambda_h = s->per_basis ? 1.0 : s->lambda_h;
  c.tolerance = s->tolerance;
  c.cost = s->cost;
  c.stop = s->stop;
  c.ab_scale = s->ab_scale;
  c.stop_le = s->lnmf ? 1 : 0;
  cost_kernel<<<1, 256, 0, h->stream>>>(c);
  return check_launch(h, "cost");
}

// Sum the per-rank partials of everything the W step and the cost need.
static int allreduce_w_inputs(nmfb_handle* h, NmfSession* s, bool with_gram) {
  if (comm_size(h->comm) <= 1) return NMFB_OK;
  const size_t nA = static_cast<size_t>(s->Kp) * s->ldw;
  const size_t nG = with_gram ? static_cast<size_t>(s->Kp) * s->Kp : 0;
  const bool kl = s->divergence == NMFB_DIV_KL;  // hs is only formed (per iteration) by the KL path
  // The m x K numerator partials travel only when every rank repeats the whole W step; with the
  // row-sharded W step (or a fixed W) each rank fetches just its rows itself (w_shard.cuh), and what is
  // left for the all-reduce are the K x K Gram matrix and the scalar sums.
  const bool small = s->w_sharded || s->W_fixed;
  NMFB_TRY(comm_allreduce(h, small ? (nG ? s->packed + nA : nullptr) : s->packed, small ? nG : nA + nG,
                          kl ? s->hs : nullptr, kl ? s->Kp : 0, s->scal, 4));
  if (with_gram && s->direct_cost) {  // otherwise run_gram_post_allreduce makes the tf32 copy
    const int cnt = s->Kp * s->Kp;
    round_copy_kernel<<<dim3((cnt + 255) / 256, 1), 256, 0, h->stream>>>(s->gramH.g32, s->gramH.gtf, 1, cnt,
                                                                       cnt, s->stop);
    NMFB_TRY(check_launch(h, "round_copy(G_H)"));
  }
  return NMFB_OK;
}

// The whole W step on this rank's rows, partial sums fetched from / results delivered to the peers
static int enqueue_w_sharded(nmfb_handle* h, NmfSession* s, int mode, bool b_partial) {
  WShardArgs a{};
  a.open_barrier = s->ws_open_barrier ? 1 : 0;
  if (!comm_peer_table(h, &a.t)) return h->fail(NMFB_ERR_CUDA, "internal: sharded W step without a peer mapping");
  a.mode = mode;
  a.K = s->K;
  a.r0 = s->r0;
  a.mb = s->mb;
  a.ld = s->ldw;
  a.a_off = s->a_off;
  a.b_off = b_partial ? s->b_off : 0;
  a.Bloc = (mode == WSTEP_EUCLID && !b_partial) ? s->B : nullptr;
  a.Wm = s->Wm;
  a.wt_off = s->wt_off;
  a.x_off = s->x_off;
  a.wsum = s->wsum;
  a.hs = s->hs;
  a.lambda = s->lambda_w;
  a.lambda_k = s->lamW_k;
  a.fixed_k = s->fixW_k;
  a.expo = s->expo;
  a.stop = s->stop;
  // TMA bulk copies for the peer traffic when one column of W fits the staging buffer (and only the numerator
  // is partial); NMFB_W_BULK=0 keeps the load / store path
  const size_t smem = static_cast<size_t>(a.t.nranks) * a.mb * sizeof(float);
  const char* env = std::getenv("NMFB_W_BULK");
  const bool bulk = !b_partial && smem <= 96 * 1024 && !(env && env[0] == '0');
  const int cap = w_shard_capacity(a.t.nranks, h->num_sms, bulk, smem);  // blocks spin on their peers: all resident
  a.rounds = (s->K + cap - 1) / cap;
  const int grid = (s->K + a.rounds - 1) / a.rounds;
  a.epoch0 = comm_next_epoch(h, a.rounds);
  a.timing = s->ws_timing;
  launch_w_step_sharded(a, grid, bulk, smem, h->stream);
  s->ws_grid = grid;
  return check_launch(h, "w_step_sharded");
}

static int enqueue_w_finish(nmfb_handle* h, NmfSession* s, int mode) {
  if (s->w_sharded) return enqueue_w_sharded(h, s, mode, false);
  WStepArgs w{};
  w.mode = mode;
  w.W = s->Wm;
  w.Wt = s->Wt;
  w.A = s->A;
  w.B = mode == WSTEP_EUCLID ? s->B : nullptr;
  w.m = s->m;
  w.ld = s->ldw;
  w.K = s->K;
  w.T = 1;
  w.cnmf_style = 0;
  w.wsum = s->wsum;
  w.hs = s->hs;
  w.lambda = s->lambda_w;
  w.stop = s->stop;
  w.lambda_k = s->lamW_k;
  w.fixed_k = s->fixW_k;
  return launch_w_step(h, w);
}

static void fill_cost_args(NmfSession* s, CostArgs* c, int iter, int mode) {
  c->mode = mode;
  c->iter = iter;
  c->Kp = s->Kp;
  c->GW = s->gramW.g32;
  c->GH = s->gramH.g32;
  c->vsq = s->vsq;
  c->vstats = s->vstats;
  c->scal = s->scal;
  c->wsum = s->wsum;
  c->n_wsum = s->Kp;
  c->lambda_w = s->per_basis ? 1.0 : s->lambda_w;
  c->lambda_h = s->per_basis ? 1.0 : s->lambda_h;
  c->lamw_k = s->lamW_k;
  c->tolerance = s->tolerance;
  c->cost = s->cost;
  c->stop = s->stop;
}

static int run_timed(nmfb_handle* h, const GemmOp& op, int which) {
  NMFB_TRY(prof_mark(h, which));
  NMFB_TRY(run_gemm(h, op));
  return prof_mark(h, which);
}

extern "C" int nmfb_profile_enable(nmfb_handle* h, int on) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  for (int w = 0; w < 5; ++w) {
    for (cudaEvent_t e : h->prof_ev[w]) cudaEventDestroy(e);
    h->prof_ev[w].clear();
  }
  h->profile = on != 0;
  return NMFB_OK;
}

// Average device time (ms) of the W-step and H-step contractions recorded so far.
extern "C" int nmfb_profile_get_all(nmfb_handle* h, double* ms_out /* [5] */) {
  if (!h || !ms_out) return NMFB_ERR_INVALID_ARGUMENT;
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int w = 0; w < 5; ++w) {
    // Launch groups whose kernels were switched off by a device-side guard (stop flag, line-search
    // phase) return in a few microseconds; they are not launches of the kernel being measured, so
    // samples below 30 % of the 90th percentile are left out of the average.
    std::vector<float> v;
    float mx = 0.f;
    for (size_t i = 0; i + 1 < h->prof_ev[w].size(); i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->prof_ev[w][i], h->prof_ev[w][i + 1]) == cudaSuccess) {
        v.push_back(ms);
        mx = std::max(mx, ms);
      }
    }
    if (!v.empty()) {  // reference = 90th percentile (one slow outlier must not hide the real launches)
      std::vector<float> sorted(v);
      std::sort(sorted.begin(), sorted.end());
      mx = sorted[(sorted.size() - 1) * 9 / 10];
    }
    double tot = 0.0;
    int c = 0;
    for (float ms : v)
      if (ms >= 0.3f * mx) {
        tot += ms;
        ++c;
      }
    ms_out[w] = c ? tot / c : 0.0;
  }
  return NMFB_OK;
}

extern "C" int nmfb_profile_get(nmfb_handle* h, double* ms_w, double* ms_h, int* count) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  double out[2] = {0.0, 0.0};
  int cnt = 0;
  for (int w = 0; w < 2; ++w) {
    double tot = 0.0;
    int c = 0;
    for (size_t i = 0; i + 1 < h->prof_ev[w].size(); i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->prof_ev[w][i], h->prof_ev[w][i + 1]) == cudaSuccess) {
        tot += ms;
        ++c;
      }
    }
    out[w] = c ? tot / c : 0.0;
    cnt = std::max(cnt, c);
  }
  if (ms_w) *ms_w = out[0];
  if (ms_h) *ms_h = out[1];
  if (count) *count = cnt;
  return NMFB_OK;
}

// Z step + H = Z*A of constrainednmf.m:213-237 from N (possibly as split slabs) and D (matrix or per-basis value)
static int enqueue_tied_update(nmfb_handle* h, NmfSession* s, const float* Nparts, int splits, long long slab,
                               long long ldn, const float* D, const float* dvec, float expo) {
  const long long pairs = static_cast<long long>(s->K) * s->nz;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(h->num_sms * 8, (pairs + 7) / 8)));
  tied_update_kernel<<<blocks, 256, 0, h->stream>>>(Nparts, splits, slab, ldn, D, dvec, s->Zm, s->ldz, s->Hm, s->Ht,
                                                    s->ldh, s->seg, s->nz, s->K, s->lambda_h, s->H_fixed ? 1 : 0, expo,
                                                    s->scal, s->stop);
  return check_launch(h, "tied_update");
}

static int enqueue_iteration_two_weight(nmfb_handle* h, NmfSession* s, int i) {
  const int K = s->K, n = s->n;
  const int cost_mode = s->divergence == NMFB_DIV_IS ? 4 : 5;
  const bool multi = comm_size(h->comm) > 1;
  auto reduce_pair = [&](const AbOp& op, float* out_n, float* out_p) {
    const long long cnt = op.args.slab;
    const int blocks = static_cast<int>(std::min<long long>((cnt + 255) / 256, 2048));
    split_reduce2_kernel<<<blocks, 256, 0, h->stream>>>(op.parts_n, op.parts_p, op.splits, cnt, out_n, out_p, cnt, s->stop);
    return check_launch(h, "split_reduce2(OUTn, OUTp)");
  };
  if (s->ab_fused) {
    // A = Qn H', B = Qp H' (+ the divergence of iteration i-1) in one fused kernel
    if (!s->W_fixed || i > 0) {
      s->abW.args.want_cost = i > 0 ? 1 : 0;
      NMFB_TRY(prof_mark(h, 0));
      NMFB_TRY(run_ab(h, s->abW));
      NMFB_TRY(prof_mark(h, 0));
      if (!s->W_fixed) NMFB_TRY(reduce_pair(s->abW, s->A, s->B));
    }
  } else {
  s->gemmS.L.args.want_cost = i > 0 ? 1 : 0;
  NMFB_TRY(run_gemm(h, s->gemmS));  // weights from the current V_hat (+ divergence of iteration i-1)
  if (!s->W_fixed) {
    NMFB_TRY(run_gemm(h, s->gemmR));
    NMFB_TRY(run_gemm(h, s->gemmRb));
  }
  }
  if (multi) {  // column shards: A, B and the cost sums are partial (one all-reduce, as for euclidean / KL)
    const bool small = s->w_sharded || s->W_fixed;  // the m x K partials are fetched row block by row block instead
    NMFB_TRY(comm_allreduce(h, small ? nullptr : s->packed, small ? 0 : 2 * static_cast<size_t>(s->Kp) * s->ldw, nullptr,
                            0, s->scal, 4));
  }
  if (i > 0) NMFB_TRY(enqueue_cost(h, s, i - 1, cost_mode));
  if (!s->W_fixed && s->w_sharded) {
    NMFB_TRY(enqueue_w_sharded(h, s, WSTEP_EUCLID, true));
    if (!s->ab_fused) {
      s->gemmS.L.args.want_cost = 0;
      NMFB_TRY(run_gemm(h, s->gemmS));  // refreshed V_hat (nmf.m:173)
    }
  } else if (!s->W_fixed) {
    WStepArgs w{};
    w.mode = WSTEP_EUCLID;  // same shape: neg = A + W diag(<W,B>), pos = B + W diag(<W,A>)
    w.W = s->Wm;
    w.Wt = s->Wt;
    w.A = s->A;
    w.B = s->B;
    w.m = s->m;
    w.ld = s->ldw;
    w.K = s->K;
    w.T = 1;
    w.wsum = s->wsum;
    w.hs = s->hs;
    w.lambda = s->lambda_w;
    w.stop = s->stop;
    w.expo = s->expo;
    w.lambda_k = s->lamW_k;
    w.fixed_k = s->fixW_k;
    NMFB_TRY(launch_w_step(h, w));
    if (!s->ab_fused) {
      s->gemmS.L.args.want_cost = 0;
      NMFB_TRY(run_gemm(h, s->gemmS));  // refreshed V_hat (nmf.m:173)
    }
  }
  if (s->ab_fused) {  // N = W' Qn, D = W' Qp from the refreshed V_hat (nmf.m:173), again without leaving the chip
    NMFB_TRY(prof_mark(h, 1));
    NMFB_TRY(run_ab(h, s->abH));
    NMFB_TRY(prof_mark(h, 1));
    if (s->tied) {
      NMFB_TRY(reduce_pair(s->abH, s->Nbuf, s->Dbuf));
    } else {  // the H update sums the partial slabs itself
      h_finish_kernel<<<dim3(std::max(1, std::min(64, (n + 1023) / 1024)), K), 256, 0, h->stream>>>(
          s->abH.parts_n, s->abH.parts_p, s->Hm, s->Ht, s->ldh, n, s->lambda_h, s->H_fixed ? 1 : 0, s->scal, s->stop,
          s->expo, s->lamH_k, s->fixH_k, s->abH.splits, s->abH.args.slab);
      return check_launch(h, "h_finish(slabs)");
    }
  } else {
  NMFB_TRY(run_gemm(h, s->gemmHn));
  NMFB_TRY(run_gemm(h, s->gemmHd));
  }
  if (s->tied) return enqueue_tied_update(h, s, s->Nbuf, 1, 0, s->ldh, s->Dbuf, nullptr, s->expo);
  h_finish_kernel<<<dim3(std::max(1, std::min(64, (n + 1023) / 1024)), K), 256, 0, h->stream>>>(
      s->Nbuf, s->Dbuf, s->Hm, s->Ht, s->ldh, n, s->lambda_h, s->H_fixed ? 1 : 0, s->scal, s->stop, s->expo, s->lamH_k,
      s->fixH_k);
  return check_launch(h, "h_finish");
}

static int enqueue_iteration(nmfb_handle* h, NmfSession* s, int i) {
  if (s->two_weight) return enqueue_iteration_two_weight(h, s, i);
  const int Kp = s->Kp, K = s->K, n = s->n;
  const int* stop = s->stop;
  const bool multi = comm_size(h->comm) > 1;
  if (s->divergence == NMFB_DIV_EUCLIDEAN) {
    const bool fused_cost = !multi && !s->direct_cost;  // reduce + <G_W,G_H> + cost in one kernel
    if (fused_cost && s->overlap) {
      // side stream: G_H, <G_W,G_H>, cost(i-1), stop test; publishes gate[0] = i+1 when G_H is ready
      CostArgs c{};
      fill_cost_args(s, &c, i - 1, 0);
      NMFB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
      NMFB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
      cudaStream_t main_stream = h->stream;
      h->stream = h->stream2;
      int rc = run_gram_cost(h, s->gramH, s->ticket, c, i > 0, s->gates + 0, static_cast<unsigned>(i + 1));
      h->stream = main_stream;
      NMFB_TRY(rc);
      s->gemmA.L.args.gate_value = static_cast<unsigned>(i + 1);
    } else if (fused_cost) {
      CostArgs c{};
      fill_cost_args(s, &c, i - 1, 0);
      NMFB_TRY(prof_mark(h, 2));
      NMFB_TRY(run_gram_cost(h, s->gramH, s->ticket, c, i > 0));
      NMFB_TRY(prof_mark(h, 2));
    } else if (s->side_gh) {
      NMFB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
      NMFB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
      cudaStream_t main_stream = h->stream;
      h->stream = h->stream2;
      int rc = run_gram(h, s->gramH, stop);
      if (rc == NMFB_OK && s->ws_open_barrier) {
        // row-sharded W step: what is left to all-reduce (G_H, the scalar sums) and the cost / stop test of the
        // previous iteration also run beside the A GEMM
        rc = allreduce_w_inputs(h, s, true);
        if (rc == NMFB_OK) {
          CostArgs c{};
          fill_cost_args(s, &c, i - 1, 0);
          rc = run_gram_post_allreduce(h, s->gramH, s->ticket, c, i > 0);
        }
      }
      h->stream = main_stream;
      NMFB_TRY(rc);
      NMFB_CUDA(h, cudaEventRecord(h->ev_join, h->stream2));
    } else if (!s->H_fixed || i == 0 || multi) {
      NMFB_TRY(run_gram(h, s->gramH, stop));
    }
    if (!s->W_fixed) NMFB_TRY(run_timed(h, s->gemmA, 0));
    if (s->side_gh) NMFB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    if (s->side_gh && s->ws_open_barrier) {
      // (all-reduce and cost already queued on the side stream)
    } else {
    NMFB_TRY(allreduce_w_inputs(h, s, true));
    if (multi && !s->direct_cost) {
      CostArgs c{};
      fill_cost_args(s, &c, i - 1, 0);
      NMFB_TRY(run_gram_post_allreduce(h, s->gramH, s->ticket, c, i > 0));
    } else if (i > 0 && !s->direct_cost && !fused_cost) {
      NMFB_TRY(enqueue_cost(h, s, i - 1, 0));
    }
    }
    if (!s->W_fixed) {
      if (s->gemmB.planned) NMFB_TRY(run_gemm(h, s->gemmB));
      NMFB_TRY(prof_mark(h, 3));
      NMFB_TRY(enqueue_w_finish(h, s, WSTEP_EUCLID));
      NMFB_TRY(prof_mark(h, 3));
      if (s->gate_h) {
        // side stream: G_W of the new W; the H-step contraction starts at once and waits at gate[1]
        NMFB_CUDA(h, cudaEventRecord(h->ev_join, h->stream));
        NMFB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_join, 0));
        cudaStream_t main_stream = h->stream;
        h->stream = h->stream2;
        int rc = run_gram(h, s->gramW, stop, s->ticket + 1, s->gates + 1, static_cast<unsigned>(i + 1));
        h->stream = main_stream;
        NMFB_TRY(rc);
        s->gemmH.L.args.gate_value = static_cast<unsigned>(i + 1);
      } else {
        NMFB_TRY(prof_mark(h, 4));
        NMFB_TRY(run_gram(h, s->gramW, stop));
        NMFB_TRY(prof_mark(h, 4));
      }
    }
    if (s->h_split) {
      NMFB_TRY(prof_mark(h, 1));
      NMFB_TRY(run_gemm(h, s->gemmH));  // N (split-K, summed) and D = G_W H
      if (s->tied) {
        NMFB_TRY(enqueue_tied_update(h, s, s->Nbuf, 1, 0, s->ldh, s->Dbuf, nullptr, 0.f));
      } else {
        h_finish_kernel<<<dim3(std::max(1, std::min(64, (n + 1023) / 1024)), K), 256, 0, h->stream>>>(
            s->Nbuf, s->Dbuf, s->Hm, s->Ht, s->ldh, n, s->lambda_h, s->H_fixed ? 1 : 0, s->scal, stop, 0.f, s->lamH_k,
            s->fixH_k);
        NMFB_TRY(check_launch(h, "h_finish"));
      }
      NMFB_TRY(prof_mark(h, 1));
    } else {
      NMFB_TRY(run_timed(h, s->gemmH, 1));
    }
    if (s->direct_cost) {
      NMFB_TRY(run_gemm(h, s->gemmS));
      if (multi) NMFB_TRY(comm_allreduce(h, nullptr, 0, nullptr, 0, s->scal, 4));
      NMFB_TRY(enqueue_cost(h, s, i, 1));
    }
  } else {
    if (!s->H_fixed || i == 0 || multi) {
      NMFB_TRY(zero_async(h, s->hs, Kp * sizeof(double)));
      vec_sums_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, K, n, s->ldh, s->hs, nullptr, stop);
      NMFB_TRY(check_launch(h, "vec_sums(H)"));
    }
    if (s->kl_fused) {
      // R = (V ./ (W H)) H' (+ the cost sums of the previous iteration) in one fused kernel
      if (!s->W_fixed || i > 0) {
        s->klW.args.want_cost = i > 0 ? 1 : 0;
        NMFB_TRY(prof_mark(h, 0));
        NMFB_TRY(run_kl(h, s->klW));
        NMFB_TRY(prof_mark(h, 0));
        if (!s->W_fixed) {
          const long long cnt = s->klW.args.slab;
          split_reduce_kernel<<<static_cast<int>(std::min<long long>((cnt + 255) / 256, 2048)), 256, 0, h->stream>>>(
              s->klW.parts, s->klW.splits, cnt, s->A, cnt, stop);
          NMFB_TRY(check_launch(h, "split_reduce(R)"));
        }
      }
    } else {
      s->gemmS.L.args.want_cost = i > 0 ? 1 : 0;
      NMFB_TRY(run_gemm(h, s->gemmS));  // Q = V ./ (W H) with the H of the previous iteration
      if (!s->W_fixed) NMFB_TRY(run_gemm(h, s->gemmR));
    }
    NMFB_TRY(allreduce_w_inputs(h, s, false));
    if (i > 0) NMFB_TRY(enqueue_cost(h, s, i - 1, 2));
    if (!s->W_fixed) {
      NMFB_TRY(enqueue_w_finish(h, s, s->lnmf ? WSTEP_LNMF : WSTEP_KL));
      D2FArgs da{s->wsum, s->wsf, Kp};
      d2f_kernel<<<(Kp + 127) / 128, 128, 0, h->stream>>>(da, stop);
      NMFB_TRY(check_launch(h, "d2f(ws)"));
      if (!s->kl_fused) {
        s->gemmS.L.args.want_cost = 0;
        NMFB_TRY(run_gemm(h, s->gemmS));  // refreshed V_hat (nmf.m:173)
      }
    }
    if (s->kl_fused) {
      // N' = (V ./ (W H))' W with the refreshed V_hat, then the H update on the summed slabs
      NMFB_TRY(prof_mark(h, 1));
      NMFB_TRY(run_kl(h, s->klH));
      NMFB_TRY(prof_mark(h, 1));
      if (s->tied) {
        NMFB_TRY(enqueue_tied_update(h, s, s->klH.parts, s->klH.splits, s->klH.args.slab, s->klH.args.ldo, nullptr,
                                     s->wsf, 0.f));
      } else {
        kl_h_finish_kernel<<<dim3(std::max(1, std::min(64, (n + 1023) / 1024)), K), 256, 0, h->stream>>>(
            s->klH.parts, s->klH.splits, s->klH.args.slab, s->klH.args.ldo, s->Hm, s->Ht, s->ldh, s->wsf, s->lambda_h, n,
            s->H_fixed ? 1 : 0, s->scal, stop, s->lamH_k, s->fixH_k, s->lnmf ? 1 : 0);
        NMFB_TRY(check_launch(h, "kl_h_finish"));
      }
    } else {
      NMFB_TRY(run_gemm(h, s->gemmH));
      if (s->kl_store_n && s->tied) {
        NMFB_TRY(enqueue_tied_update(h, s, s->Nbuf, 1, 0, s->ldh, nullptr, s->wsf, 0.f));
      } else if (s->kl_store_n) {
        kl_h_finish_kernel<<<dim3(std::max(1, std::min(64, (n + 1023) / 1024)), K), 256, 0, h->stream>>>(
            s->Nbuf, 1, 0, s->ldh, s->Hm, s->Ht, s->ldh, s->wsf, s->lambda_h, n, s->H_fixed ? 1 : 0, s->scal, stop,
            s->lamH_k, s->fixH_k, s->lnmf ? 1 : 0);
        NMFB_TRY(check_launch(h, "kl_h_finish"));
      }
    }
  }
  return NMFB_OK;
}

// cost of the last executed iteration (needs quantities of the final H)
static int enqueue_final_cost(nmfb_handle* h, NmfSession* s) {
  if (s->finalized || s->iters_enqueued == 0) return NMFB_OK;
  s->finalized = true;
  const int last = s->iters_enqueued - 1;
  const bool multi = comm_size(h->comm) > 1;
  if (s->two_weight) {
    if (s->ab_fused) {
      s->abW.args.want_cost = 1;
      NMFB_TRY(run_ab(h, s->abW));
    } else {
      s->gemmS.L.args.want_cost = 1;
      NMFB_TRY(run_gemm(h, s->gemmS));
    }
    if (multi) NMFB_TRY(comm_allreduce(h, nullptr, 0, nullptr, 0, s->scal, 4));
    return enqueue_cost(h, s, last, s->divergence == NMFB_DIV_IS ? 4 : 5);
  }
  if (s->divergence == NMFB_DIV_EUCLIDEAN) {
    if (s->direct_cost) return NMFB_OK;
    if (!multi) {
      CostArgs c{};
      fill_cost_args(s, &c, last, 0);
      return run_gram_cost(h, s->gramH, s->ticket, c, true);
    }
    NMFB_TRY(run_gram(h, s->gramH, s->stop));
    NMFB_TRY(comm_allreduce(h, s->gramH.g32, static_cast<size_t>(s->Kp) * s->Kp, nullptr, 0, s->scal, 4));
    CostArgs c{};
    fill_cost_args(s, &c, last, 0);
    return run_gram_post_allreduce(h, s->gramH, s->ticket, c, true);
  }
  if (s->kl_fused) {
    s->klW.args.want_cost = 1;
    NMFB_TRY(run_kl(h, s->klW));
  } else {
    s->gemmS.L.args.want_cost = 1;
    NMFB_TRY(run_gemm(h, s->gemmS));
  }
  if (multi) NMFB_TRY(comm_allreduce(h, nullptr, 0, nullptr, 0, s->scal, 4));
  return enqueue_cost(h, s, last, 2);
}

void nmf_session_release(nmfb_handle* h) {
  if (h->sess) {
    delete h->sess;
    h->sess = nullptr;
  }
}

// ------------------------------------------------------------------ C ABI
static int nmf_begin_impl(nmfb_handle* h, int K, const nmfb_config* cfg, bool lnmf) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  nmf_session_release(h);
  h->sess = new NmfSession();
  int rc = nmf_setup(h, h->sess, K, cfg, lnmf);
  if (rc == NMFB_OK) {
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = h->fail(NMFB_ERR_CUDA, "nmf setup: %s", cudaGetErrorString(e));
  }
  if (rc != NMFB_OK) nmf_session_release(h);
  return rc;
}
extern "C" int nmfb_nmf_begin(nmfb_handle* h, int K, const nmfb_config* cfg) {
  return nmf_begin_impl(h, K, cfg, false);
}

extern "C" int nmfb_nmf_step(nmfb_handle* h, int iters) {
  if (!h || !h->sess) return NMFB_ERR_INVALID_ARGUMENT;
  NmfSession* s = h->sess;
  if (s->finalized) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "nmf_step after the cost was finalised");
  cudaSetDevice(h->device);
  iters = std::min(iters, s->maxiter - s->iters_enqueued);
  if (iters <= 0) return NMFB_OK;
  NMFB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  for (int k = 0; k < iters; ++k) {
    NMFB_TRY(enqueue_iteration(h, s, s->iters_enqueued));
    ++s->iters_enqueued;
  }
  NMFB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  NMFB_CUDA(h, cudaMemcpyAsync(s->pinned, s->stop, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return NMFB_OK;
}

extern "C" int nmfb_nmf_sync(nmfb_handle* h, int* iters_done, double* device_ms) {
  if (!h || !h->sess) return NMFB_ERR_INVALID_ARGUMENT;
  NmfSession* s = h->sess;
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  if (s->iters_enqueued > 0 && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess)
    s->device_ms = ms;
  if (iters_done) *iters_done = s->iters_enqueued;
  if (device_ms) *device_ms = s->device_ms;
  return NMFB_OK;
}

static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

extern "C" int nmfb_nmf_end(nmfb_handle* h, float* W_out, float* H_out, double* cost_out, int* n_cost) {
  if (!h || !h->sess) return NMFB_ERR_INVALID_ARGUMENT;
  NmfSession* s = h->sess;
  cudaSetDevice(h->device);
  const bool trace = std::getenv("NMFB_TRACE") != nullptr;
  const double t0 = now_ms();
  int rc = enqueue_final_cost(h, s);
  if (rc == NMFB_OK) {
    cudaError_t e = cudaMemcpyAsync(s->pinned, s->stop, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = h->fail(NMFB_ERR_CUDA, "nmf_end: %s", cudaGetErrorString(e));
  }
  if (rc == NMFB_OK) {
    // lnmf.m:88-91 leaves the loop WITHOUT trimming: cost keeps maxiter entries, zeros after the stop
    const int nc = s->lnmf ? s->maxiter : s->pinned[1];
    if (n_cost) *n_cost = nc;
    if (cost_out && nc > 0) {
      cudaError_t e = cudaMemcpy(cost_out, s->cost, nc * sizeof(double), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = h->fail(NMFB_ERR_CUDA, "cost download: %s", cudaGetErrorString(e));
    }
  }
  if (rc == NMFB_OK && s->ws_timing != nullptr && s->ws_grid > 0) {
    std::vector<unsigned long long> t(8 * static_cast<size_t>(s->ws_grid));
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(t.data(), s->ws_timing, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t6 = 0;
    double ph[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < s->ws_grid; ++b) {
      t0 = std::min(t0, t[b * 8]);
      t6 = std::max(t6, t[b * 8 + 6]);
      for (int i = 0; i < 6; ++i) ph[i] += static_cast<double>(t[b * 8 + i + 1] - t[b * 8 + i]) / s->ws_grid;
    }
    fprintf(stderr, "[nmfb] sharded W step, rank %d, last launch: %d blocks, kernel span %.1f us; mean per block (us): peer rows + dots %.1f, "
                    "barrier %.1f, step + norms %.1f, barrier %.1f, normalise + deliver rows %.1f, closing barrier %.1f\n",
            comm_rank(h->comm), s->ws_grid, (t6 - t0) * 1e-3, ph[0] * 1e-3, ph[1] * 1e-3, ph[2] * 1e-3, ph[3] * 1e-3, ph[4] * 1e-3, ph[5] * 1e-3);
  }
  if (rc == NMFB_OK && s->w_sharded && !s->W_fixed) {
    // every rank kept only its rows of the fp32 W current: one exchange makes W whole everywhere
    // (a collective: issued by every rank whether or not it asked for W)
    WGatherArgs g{};
    if (comm_peer_table(h, &g.t)) {
      g.K = s->K;
      g.r0 = s->r0;
      g.mb = s->mb;
      g.ld = s->ldw;
      g.wm_off = s->wm_off;
      g.epoch = comm_next_epoch(h, 1);
      w_gather_rows_kernel<<<std::min(s->K, std::min(h->num_sms, kMaxBlocks)), kWsThreads, 0, h->stream>>>(g);
      rc = check_launch(h, "w_gather_rows");
      if (rc == NMFB_OK && cudaStreamSynchronize(h->stream) != cudaSuccess)
        rc = h->fail(NMFB_ERR_CUDA, "gather of the W rows failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  const double t1 = now_ms();
  if (rc == NMFB_OK && W_out) rc = download_colmajor(h, s->Wm, s->ldw, s->m, s->K, W_out);
  const double t2 = now_ms();
  if (rc == NMFB_OK && H_out) rc = download_H(h, s->Hm, s->ldh, s->K, s->n, H_out);
  const double t3 = now_ms();
  nmf_session_release(h);
  if (trace)
    fprintf(stderr, "[nmfb] nmf_end: final cost + trace %.1f ms, W download %.1f ms, H download %.1f ms, release %.1f ms, "
                    "%lld cudaMalloc so far\n", t1 - t0, t2 - t1, t3 - t2, now_ms() - t3, h->mallocs);
  return rc;
}

static int nmf_call(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out, double* cost_out,
                    int* n_cost, bool lnmf) {
  const bool trace = std::getenv("NMFB_TRACE") != nullptr;
  const double t0 = now_ms();
  NMFB_TRY(nmf_begin_impl(h, K, cfg, lnmf));
  const double t1 = now_ms();
  NmfSession* s = h->sess;
  loop_begin(h);
  int rc = run_chunked(h, s->maxiter, s->stop, [&](int i) {
    int r = enqueue_iteration(h, s, i);
    if (r == NMFB_OK) ++s->iters_enqueued;
    return r;
  });
  if (rc != NMFB_OK) {
    nmf_session_release(h);
    return rc;
  }
  const double t2 = now_ms();
  // drain both streams before the (allocating, copying) epilogue of the call: measured to avoid
  // a pathological slow path of cudaMalloc / pageable copies issued under a deep launch queue
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->stream2);
  loop_end(h, s->iters_enqueued);
  const double t3 = now_ms();
  int nc_local = 0;
  rc = nmfb_nmf_end(h, W_out, H_out, cost_out, n_cost ? n_cost : &nc_local);
  h->loop_iters = lnmf ? h->loop_iters : (n_cost ? *n_cost : nc_local);  // executed (not merely queued) iterations
  if (trace)
    fprintf(stderr, "[nmfb] nmf: setup %.1f ms, enqueue %.1f ms, drain %.1f ms, finish %.1f ms\n", t1 - t0, t2 - t1,
            t3 - t2, now_ms() - t3);
  return rc;
}

extern "C" int nmfb_nmf(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out,
                        double* cost_out, int* n_cost) {
  return nmf_call(h, K, cfg, W_out, H_out, cost_out, n_cost, false);
}

// lnmf.m:1 - local NMF: KL-type multiplicative updates with unit-sum bases (lnmf.m:63,75) and the
// square-root H step (lnmf.m:81), on the fused KL kernels.
extern "C" int nmfb_lnmf(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out,
                         double* cost_out, int* n_cost) {
  return nmf_call(h, K, cfg, W_out, H_out, cost_out, n_cost, true);
}

// constrainednmf.m:1 - label-constrained NMF V ~ W*Z*A (Liu & Wu): the W step is nmf.m's, the encoding is
// H = Z*A with A the label-indicator matrix, and Z takes the multiplicative step on the class-summed
// gradients (constrainednmf.m:213-237).  The caller orders the samples as the reference does (unlabeled
// first, classes contiguous; constrainednmf.m:147-164) and passes the column map of that arrangement.
extern "C" int nmfb_constrainednmf(nmfb_handle* h, int K, const nmfb_config* cfg, const int* col2z, int nz,
                                   const float* Z_init, float* W_out, float* H_out, float* Z_out, double* cost_out,
                                   int* n_cost) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  if (!col2z || nz <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "constrainednmf: column map missing");
  if (cfg && cfg->divergence == NMFB_DIV_AB && cfg->alpha != 0 && h->m != K)
    // constrainednmf.m:229 multiplies a K x n by an m x n matrix element-wise in the non-dual AB branch
    return h->fail(NMFB_ERR_UNSUPPORTED, "Matrix dimensions must agree. (constrainednmf.m:229: the non-dual "
                                         "alpha-beta Z update of the reference is only defined for m == num_basis_elems)");
  h->tie_col2z = col2z;
  h->tie_nz = nz;
  h->tie_Zinit = Z_init;
  const double t0 = now_ms();
  int rc = nmf_begin_impl(h, K, cfg, false);
  h->tie_col2z = nullptr;
  h->tie_Zinit = nullptr;
  if (rc != NMFB_OK) return rc;
  (void)t0;
  NmfSession* s = h->sess;
  loop_begin(h);
  rc = run_chunked(h, s->maxiter, s->stop, [&](int i) {
    int r = enqueue_iteration(h, s, i);
    if (r == NMFB_OK) ++s->iters_enqueued;
    return r;
  });
  if (rc != NMFB_OK) {
    nmf_session_release(h);
    return rc;
  }
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->stream2);
  loop_end(h, s->iters_enqueued);
  if (Z_out) {
    rc = download_H(h, s->Zm, s->ldz, s->K, s->nz, Z_out);
    if (rc != NMFB_OK) {
      nmf_session_release(h);
      return rc;
    }
  }
  int nc_local = 0;
  rc = nmfb_nmf_end(h, W_out, H_out, cost_out, n_cost ? n_cost : &nc_local);
  h->loop_iters = n_cost ? *n_cost : nc_local;
  return rc;
}
