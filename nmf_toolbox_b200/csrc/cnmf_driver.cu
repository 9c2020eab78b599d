// cnmf driver: the iteration loop of cnmf.m (lines 175-258) in stacked form.  Euclidean /
// 'frobenius' run on the Gram path described here; KL, IS and AB (which cnmf.m treats as one
// alpha-beta family, cnmf.m:137-147) on the two-weight path further down.
//
// With Wc = [W_1 ... W_T] (m x KT, which is exactly the column-major memory of
// the reference's m x K x T tensor) and Hs = [H_1; ...; H_T] (KT x n, H_t = H
// shifted right by t-1 columns with zero fill, cnmf.m:188) the reconstruction is
// V_hat = Wc*Hs (ReconstructFromDecomposition.m:33-38) and
//   - the T per-frame W updates of cnmf.m:187-194 (all computed from the same
//     stale V_hat and H) are ONE nmf.m-style update of Wc with
//     A = V*Hs', B = Wc*(Hs*Hs') and per-column dots <Wc_c, A_c>, <Wc_c, B_c>;
//   - the H numerator / denominator of cnmf.m:218-227 are fold(Wc'*V) and
//     fold((Wc'*Wc)*Hs) with fold(P)(k, j) = sum_t P(k + K(t-1), j + t - 1);
//   - the Euclidean cost is 0.5*(|V|^2 - 2<Wc'V, Hs> + <Wc'Wc, Hs*Hs'>).
// Basis k is normalised over all frames by |W(:,k,:)|_F / T (cnmf.m:196-199);
// H is compensated for that only at initialisation (cnmf.m:163).
//
// Several GPUs (Euclidean / 'frobenius'): V and H are sharded by columns (time), W is replicated.  The shifts of
// cnmf.m:188 (H moved right by t-1) and cnmf.m:219 (V, V_hat moved left by t-1) reach T-1 columns across a shard
// boundary, so every rank keeps a HALO: the T-1 columns of H before and after its own ones (exchanged once per
// iteration through the peer region, a few KB) and the T-1 columns of V after them (exchanged once, at setup).
// With those, Hs, P = Wc'V and D = (Wc'Wc)Hs are formed locally on own + halo columns and folded for the own
// columns; what crosses the GPUs per iteration besides the H halo is one all-reduce of [A | Hs Hs'] and the
// scalar sums, as for nmf.
#include <algorithm>
#include <cstring>
#include <vector>

#include "comm.cuh"
#include "engine.cuh"
#include "ew_kernels.cuh"

using namespace nmfb;

namespace cnmfdetail {

struct CnmfState {
  Arena ar;
  int K = 0, T = 0, KT = 0, KTp = 0, m = 0, n = 0;
  long long ldw = 0, ldh = 0;
  // column shards: own columns n, halo columns hL (left, H only) and hR (right, H and V); nx = n + hR columns
  // of Hs / P / D; the H master lives in a buffer of hL + n + hR columns (leading dimension ldhm), Hm points
  // at the first own column
  bool multi = false;
  int hL = 0, hR = 0, nx = 0, halo = 0, rank = 0, nranks = 1;
  long long ldhm = 0;
  float *Hx = nullptr, *Vx = nullptr, *packed = nullptr;
  const float* Vmma = nullptr;
  bool fold_stacks = false;   // single GPU: fold_update writes the next iteration's Hs itself
  unsigned int* ticket = nullptr;
  size_t send_off = 0;  // byte offset of this rank's [2][K][halo] H columns (first / last own ones) in the region
  GemmOp gemmB;
  bool W_fixed = false, H_fixed = false, frobenius = false;
  float lambda_w = 0.f, lambda_h = 0.f;
  int maxiter = 100;
  double tolerance = 1e-3;
  float *Wm = nullptr, *Wt = nullptr, *Hm = nullptr, *Hs = nullptr;
  float *A = nullptr, *B = nullptr, *P = nullptr, *D = nullptr;
  float *pcoef = nullptr, *qcoef = nullptr, *bvec = nullptr, *hscale = nullptr;
  double *ab = nullptr, *norm2 = nullptr, *wsum = nullptr, *scal = nullptr, *cost = nullptr;
  int* stop = nullptr;
  double vsq = 0.0;
  GramOp gramH, gramW;
  GemmOp gemmA, gemmP;
  // KL / IS / AB: element-wise weights Qn, Qp of V_hat = Wc Hs (see plan_two_weight in nmf_driver.cu)
  bool two_weight = false, kl_quirk = false;
  float *Q = nullptr, *Q2 = nullptr;
  GemmOp gemmS, gemmRa, gemmRb, gemmPn, gemmPd;
  float expo = 0.f;
  double ab_scale = 0.0;
  int cost_mode = 0;
};

int zero_async(nmfb_handle* h, void* p, size_t bytes) {
  NMFB_CUDA(h, cudaMemsetAsync(p, 0, bytes, h->stream));
  return NMFB_OK;
}

int enqueue_cost(nmfb_handle* h, CnmfState* s, int iter) {
  if (!s->frobenius && !s->two_weight) {
    const int cnt = s->KTp * s->KTp;
    gram_dot_kernel<<<std::min(64, (cnt + 1023) / 1024), 256, 0, h->stream>>>(s->gramW.g32, s->gramH.g32, cnt,
                                                                               s->scal + 4, s->stop);
    NMFB_TRY(check_launch(h, "gram_dot"));
  }
  CostArgs c{};
  c.mode = s->two_weight ? s->cost_mode : s->frobenius ? 3 : 0;  // cnmf.m:239-248 has no 'frobenius' case
  c.ab_scale = s->ab_scale;
  c.iter = iter;
  c.Kp = s->KTp;
  c.GW = s->gramW.g32;
  c.GH = s->gramH.g32;
  c.vsq = s->vsq;
  c.scal = s->scal;
  c.wsum = s->wsum;
  c.n_wsum = s->KTp;
  c.lambda_w = s->lambda_w;
  c.lambda_h = s->lambda_h;
  c.tolerance = s->tolerance;
  c.cost = s->cost;
  c.stop = s->stop;
  cost_kernel<<<1, 256, 0, h->stream>>>(c);
  return check_launch(h, "cost");
}

// cnmf.m:196-199 after the per-column W step: W(:,k,:) /= |W(:,k,:)|_F / T, tf32 copy, column sums
int enqueue_w_scale(nmfb_handle* h, CnmfState* s) {
  NMFB_TRY(zero_async(h, s->wsum, s->KTp * sizeof(double)));
  w_normalize_kernel<<<vec_grid(s->m, s->KT), 256, 0, h->stream>>>(s->Wm, s->Wt, s->m, s->ldw, s->K, s->T, 1, s->norm2,
                                                                  s->wsum, nullptr, s->stop);
  return check_launch(h, "w_normalize");
}

// H halo exchange: every rank publishes its first and its last `halo` own columns of H, meets its peers, and reads
// the left neighbour's last / the right neighbour's first columns into its halo columns.  One block.  (The slots are
// rewritten one iteration later, behind the all-reduce of this iteration, which no rank passes before every rank has
// left this kernel.)
struct HaloArgs {
  PeerTable t;
  float* Hm;         // first own column
  long long ld;
  int K, n, halo, hL, hR;
  size_t send_off;   // region byte offset of [first | last] columns, K * halo floats each
  int epoch;
  const int* stop;
};
__global__ void cnmf_halo_kernel(HaloArgs a) {
  NMFB_STOP_GUARD(a.stop);
  float* mine = reinterpret_cast<float*>(a.t.base[a.t.rank] + a.send_off);
  const int cnt = a.K * a.halo;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int k = i / a.halo, c = i % a.halo;
    mine[i] = a.Hm[k * a.ld + c];                          // first own columns -> left neighbour's right halo
    mine[cnt + i] = a.Hm[k * a.ld + (a.n - a.halo + c)];   // last own columns  -> right neighbour's left halo
  }
  p2p_block_barrier(a.t, 7 * kFlagBytes, a.epoch);
  if (a.hL > 0) {
    const float* left = reinterpret_cast<const float*>(a.t.base[a.t.rank - 1] + a.send_off) + cnt;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int k = i / a.halo, c = i % a.halo;
      a.Hm[k * a.ld - a.halo + c] = __ldcv(left + i);
    }
  }
  if (a.hR > 0) {
    const float* right = reinterpret_cast<const float*>(a.t.base[a.t.rank + 1] + a.send_off);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int k = i / a.halo, c = i % a.halo;
      if (c < a.hR) a.Hm[k * a.ld + a.n + c] = __ldcv(right + i);
    }
  }
}
int enqueue_halo(nmfb_handle* h, CnmfState* s, const int* stop) {
  if (!s->multi || s->halo == 0) return NMFB_OK;
  HaloArgs a{};
  if (!comm_peer_table(h, &a.t)) return h->fail(NMFB_ERR_CUDA, "internal: cnmf halo exchange without a peer mapping");
  a.Hm = s->Hm;
  a.ld = s->ldhm;
  a.K = s->K;
  a.n = s->n;
  a.halo = s->halo;
  a.hL = s->hL;
  a.hR = s->hR;
  a.send_off = s->send_off;
  a.epoch = comm_next_epoch(h, 1);
  a.stop = stop;
  cnmf_halo_kernel<<<1, 256, 0, h->stream>>>(a);
  return check_launch(h, "cnmf_halo");
}

int enqueue_hstack(nmfb_handle* h, CnmfState* s, const int* stop) {
  NMFB_TRY(enqueue_halo(h, s, stop));
  hstack_kernel<<<vec_grid(s->nx, s->KT), 256, 0, h->stream>>>(s->Hm, s->Hs, s->K, s->T, s->nx, s->ldhm, s->ldh, s->hL, stop);
  return check_launch(h, "hstack");
}

// KL / IS / AB (cnmf.m:177-233 with alpha, beta from cnmf.m:137-147): both gradients of every frame
// are contractions with Qn = V^a V_hat^(b-1), Qp = V_hat^(a+b-1) (dual: V^(a-1) V_hat^b, V^(a+b-1)),
// so with the stacked operands A = Qn Hs', B = Qp Hs' feed the same per-column update as the
// Euclidean path, and the H gradients are fold(Wc' Qn), fold(Wc' Qp) - except that the KL branch
// of cnmf.m:221-222 leaves V_pos unshifted.
int enqueue_iteration_two_weight(nmfb_handle* h, CnmfState* s, int i) {
  const int* stop = s->stop;
  if (i == 0 || (!s->H_fixed && !s->fold_stacks)) NMFB_TRY(enqueue_hstack(h, s, stop));
  s->gemmS.L.args.want_cost = i > 0 ? 1 : 0;
  NMFB_TRY(run_gemm(h, s->gemmS));
  if (i > 0) NMFB_TRY(enqueue_cost(h, s, i - 1));
  if (!s->W_fixed) {
    NMFB_TRY(run_gemm(h, s->gemmRa));
    NMFB_TRY(run_gemm(h, s->gemmRb));
    WStepArgs w{};
    w.mode = WSTEP_EUCLID;
    w.W = s->Wm;
    w.Wt = s->Wt;
    w.A = s->A;
    w.B = s->B;
    w.m = s->m;
    w.ld = s->ldw;
    w.K = s->KT;
    w.T = 1;
    w.cnmf_style = 2;
    w.norm2_out = s->norm2;
    w.wsum = s->wsum;
    w.lambda = s->lambda_w;
    w.stop = stop;
    w.expo = s->expo;
    NMFB_TRY(launch_w_step(h, w));
    NMFB_TRY(enqueue_w_scale(h, s));
    s->gemmS.L.args.want_cost = 0;
    NMFB_TRY(run_gemm(h, s->gemmS));  // refreshed V_hat (cnmf.m:204)
  }
  NMFB_TRY(run_gemm(h, s->gemmPn));
  NMFB_TRY(run_gemm(h, s->gemmPd));
  fold_update_kernel<<<vec_grid(s->n, s->K), 256, 0, h->stream>>>(s->P, s->D, s->Hm, s->K, s->T, s->n, s->nx, s->ldh, s->ldhm,
                                                                  s->lambda_h, s->H_fixed ? 1 : 0, s->scal, stop,
                                                                  s->expo, s->kl_quirk ? 1 : 0,
                                                                  s->fold_stacks ? s->Hs : nullptr, s->ldh);
  return check_launch(h, "fold_update");
}

int enqueue_iteration(nmfb_handle* h, CnmfState* s, int i) {
  if (s->two_weight) return enqueue_iteration_two_weight(h, s, i);
  const int* stop = s->stop;
  const int m = s->m;
  bool cost_done = false;
  if (!s->H_fixed || i == 0) {
    if (i == 0 || !s->fold_stacks) NMFB_TRY(enqueue_hstack(h, s, stop));
    if (!s->multi && !s->frobenius) {
      // slab sum of Hs Hs' fused with <Wc'Wc, Hs Hs'>, the cost of the previous iteration and the stop test
      CostArgs c{};
      c.mode = 0;
      c.iter = i - 1;
      c.Kp = s->KTp;
      c.GW = s->gramW.g32;
      c.GH = s->gramH.g32;
      c.vsq = s->vsq;
      c.scal = s->scal;
      c.wsum = s->wsum;
      c.n_wsum = s->KTp;
      c.lambda_w = s->lambda_w;
      c.lambda_h = s->lambda_h;
      c.tolerance = s->tolerance;
      c.cost = s->cost;
      c.stop = s->stop;
      NMFB_TRY(run_gram_cost(h, s->gramH, s->ticket, c, i > 0));
      cost_done = true;
    } else {
      NMFB_TRY(run_gram(h, s->gramH, stop));
    }
  }
  if (s->multi) {  // partial sums over the column shards: [A | Hs Hs'] and the scalars in one all-reduce
    if (!s->W_fixed) NMFB_TRY(run_gemm(h, s->gemmA));
    const size_t nA = static_cast<size_t>(s->KTp) * s->ldw, nG = static_cast<size_t>(s->KTp) * s->KTp;
    const bool sendA = !s->W_fixed, sendG = !s->H_fixed || i == 0;  // (a fixed H keeps its summed Gram matrix)
    float* first = sendA ? s->packed : (sendG ? s->packed + nA : nullptr);
    NMFB_TRY(comm_allreduce(h, first, (sendA ? nA : 0) + (sendG ? nG : 0), nullptr, 0, s->scal, 4));
    if (sendG) {
      const int cnt = s->KTp * s->KTp;
      round_copy_kernel<<<dim3((cnt + 255) / 256, 1), 256, 0, h->stream>>>(s->gramH.g32, s->gramH.gtf, 1, cnt, cnt, stop);
      NMFB_TRY(check_launch(h, "round_copy(G_H)"));
    }
  }
  if (i > 0 && !cost_done) NMFB_TRY(enqueue_cost(h, s, i - 1));
  if (!s->W_fixed) {
    NMFB_TRY(prof_mark(h, 0));
    if (s->multi) NMFB_TRY(run_gemm(h, s->gemmB));  // B = Wc (Hs Hs') with the summed Gram matrix
    else NMFB_TRY(run_gemm(h, s->gemmA));           // A = V Hs', B = Wc (Hs Hs')
    NMFB_TRY(prof_mark(h, 0));
    WStepArgs w{};  // dots, multiplicative step and per-basis normalisation in one launch (CTA per basis)
    w.mode = WSTEP_EUCLID;
    w.W = s->Wm;
    w.Wt = s->Wt;
    w.A = s->A;
    w.B = s->B;
    w.m = m;
    w.ld = s->ldw;
    w.K = s->KT;  // one CTA per frame-column; the per-basis scale follows in w_normalize
    w.T = 1;
    w.cnmf_style = 2;
    w.norm2_out = s->norm2;
    w.wsum = s->wsum;
    w.hs = nullptr;
    w.lambda = s->lambda_w;
    w.stop = stop;
    NMFB_TRY(launch_w_step(h, w));
    NMFB_TRY(enqueue_w_scale(h, s));
    NMFB_TRY(run_gram(h, s->gramW, stop));
  }
  // P = Wc'V and D = (Wc'Wc) Hs, then fold over the frames and update H (cnmf.m:216-231)
  NMFB_TRY(prof_mark(h, 1));
  NMFB_TRY(run_gemm(h, s->gemmP));
  NMFB_TRY(prof_mark(h, 1));
  NMFB_TRY(prof_mark(h, 2));
  fold_update_kernel<<<vec_grid(s->n, s->K), 256, 0, h->stream>>>(s->P, s->D, s->Hm, s->K, s->T, s->n, s->nx, s->ldh, s->ldhm,
                                                                  s->lambda_h, s->H_fixed ? 1 : 0, s->scal,
                                                                  stop, 0.f, 0, s->fold_stacks ? s->Hs : nullptr, s->ldh);
  NMFB_TRY(check_launch(h, "fold_update"));
  return prof_mark(h, 2);
}

int cnmf_finish(nmfb_handle* h, CnmfState* s, float* W_out, float* H_out, double* cost_out, int* n_cost) {
  int flags[2] = {0, 0};
  NMFB_CUDA(h, cudaMemcpyAsync(flags, s->stop, sizeof(flags), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  const int nc = flags[1];
  h->loop_iters = nc;  // executed iterations (launches queued behind the stop flag were no-ops)
  if (n_cost) *n_cost = nc;
  if (cost_out && nc > 0)
    NMFB_CUDA(h, cudaMemcpy(cost_out, s->cost, nc * sizeof(double), cudaMemcpyDeviceToHost));
  if (W_out) NMFB_TRY(download_colmajor(h, s->Wm, s->ldw, s->m, s->KT, W_out));
  if (H_out) NMFB_TRY(download_H(h, s->Hm, s->ldhm, s->K, s->n, H_out));
  return NMFB_OK;
}

int cnmf_run(nmfb_handle* h, CnmfState* s, int K, int T, const nmfb_config* cfg_in, float* W_out,
             float* H_out, double* cost_out, int* n_cost) {
  if (h->Vraw == nullptr) return h->fail(NMFB_ERR_NO_DATA, "cnmf: call nmfb_set_V first");
  if (K <= 0 || T <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "cnmf: K and context_len must be positive");
  s->multi = comm_size(h->comm) > 1;
  s->nranks = comm_size(h->comm);
  s->rank = comm_rank(h->comm);
  nmfb_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  if (cfg_in) cfg = *cfg_in;
  if (cfg.maxiter <= 0) cfg.maxiter = 100;         // cnmf.m:440-442
  if (!(cfg.tolerance > 0)) cfg.tolerance = 1e-3;  // cnmf.m:445-447
  if (cfg.W_sparsity < 0) cfg.W_sparsity = 0;
  if (cfg.H_sparsity < 0) cfg.H_sparsity = 0;
  switch (cfg.divergence) {
    case NMFB_DIV_EUCLIDEAN:
      break;
    case NMFB_DIV_FROBENIUS:
      s->frobenius = true;
      break;
    case NMFB_DIV_AB:
      if (cfg.alpha == 0 && cfg.beta == 0)  // cnmf.m:133-135
        return h->fail(NMFB_ERR_AB_ZERO, "alpha = 0 and beta = 0 is not supported at this time.");
      s->two_weight = true;
      s->cost_mode = 5;
      break;
    case NMFB_DIV_KL:  // cnmf.m:141-143
      cfg.alpha = 1;
      cfg.beta = 0;
      s->two_weight = s->kl_quirk = true;
      s->cost_mode = 4;
      break;
    case NMFB_DIV_IS:  // cnmf.m:144-146
      cfg.alpha = 1;
      cfg.beta = -1;
      s->two_weight = true;
      s->cost_mode = 4;
      break;
    default:
      return h->fail(NMFB_ERR_DIVERGENCE, "unknown divergence %d", cfg.divergence);
  }
  const int m = h->m, n = h->n;
  s->K = K;
  s->T = T;
  s->KT = K * T;
  s->KTp = round_up(s->KT, 32);
  s->m = m;
  s->n = n;
  s->ldw = round_up(m, 4);
  s->nx = n;
  s->ldh = s->ldhm = round_up(n, 4);
  char* region = nullptr;
  size_t vhalo_off = 0;
  if (s->multi) {
    if (s->two_weight)
      return h->fail(NMFB_ERR_UNSUPPORTED, "cnmf on several GPUs: 'euclidean' / 'frobenius' only");
    s->halo = T - 1;
    // region = [A (KTp x ldw) | Hs Hs' (KTp x KTp)] [scal 8 | shard sizes 8 doubles] [V halo: halo x ldv] [H columns 2 x K x halo]
    const size_t nA = static_cast<size_t>(s->KTp) * s->ldw, nG = static_cast<size_t>(s->KTp) * s->KTp;
    const size_t dbl_off = ((nA + nG) * sizeof(float) + 255) / 256 * 256;
    vhalo_off = dbl_off + 256;
    const size_t send_off = vhalo_off + (static_cast<size_t>(std::max(1, s->halo)) * h->ldv * sizeof(float) + 255) / 256 * 256;
    const size_t total = send_off + 2 * static_cast<size_t>(K) * std::max(1, s->halo) * sizeof(float);
    NMFB_TRY(comm_acquire_region(h, total, &region));
    if (!comm_peer_table(h, nullptr))
      return h->fail(NMFB_ERR_UNSUPPORTED, "cnmf on several GPUs needs peer access between them (CUDA IPC over NVLink)");
    s->packed = reinterpret_cast<float*>(region);
    s->scal = reinterpret_cast<double*>(region + dbl_off);
    s->send_off = comm_region_offset(h, region) + send_off;
    // every rank learns all shard widths: halos only reach the direct neighbours
    double* table = s->scal + 8;
    double mine = n;
    NMFB_CUDA(h, cudaMemcpyAsync(table + s->rank, &mine, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    NMFB_TRY(comm_allreduce(h, nullptr, 0, table, s->nranks, nullptr, 0));
    double widths[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    NMFB_CUDA(h, cudaMemcpyAsync(widths, table, s->nranks * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < s->nranks; ++q)
      if (widths[q] < s->halo)
        return h->fail(NMFB_ERR_INVALID_ARGUMENT, "cnmf on several GPUs: every shard needs at least context_len - 1 = %d "
                                                  "columns (rank %d has %d)", s->halo, q, static_cast<int>(widths[q]));
    s->hL = s->rank > 0 ? s->halo : 0;
    s->hR = s->rank + 1 < s->nranks ? s->halo : 0;
    s->nx = n + s->hR;
    s->ldh = round_up(s->nx, 4);
    s->ldhm = round_up(s->hL + n + s->hR, 4);
  }
  s->W_fixed = cfg.W_fixed != 0;
  s->H_fixed = cfg.H_fixed != 0;
  s->lambda_w = static_cast<float>(cfg.W_sparsity);
  s->lambda_h = static_cast<float>(cfg.H_sparsity);
  s->maxiter = cfg.maxiter;
  s->tolerance = cfg.tolerance;
  const int KT = s->KT, KTp = s->KTp;
  Arena* ar = &s->ar;
  NMFB_TRY(ar->alloc(h, &s->Wm, static_cast<size_t>(KTp) * s->ldw));
  NMFB_TRY(ar->alloc(h, &s->Wt, static_cast<size_t>(KTp) * s->ldw));
  NMFB_TRY(ar->alloc(h, &s->Hx, static_cast<size_t>(K) * s->ldhm));
  s->Hm = s->Hx + s->hL;
  NMFB_TRY(ar->alloc(h, &s->Hs, static_cast<size_t>(KTp) * s->ldh));
  if (s->multi) s->A = s->packed;
  else NMFB_TRY(ar->alloc(h, &s->A, static_cast<size_t>(KTp) * s->ldw));
  NMFB_TRY(ar->alloc(h, &s->B, static_cast<size_t>(KTp) * s->ldw));
  NMFB_TRY(ar->alloc(h, &s->P, static_cast<size_t>(KTp) * s->ldh));
  NMFB_TRY(ar->alloc(h, &s->D, static_cast<size_t>(KTp) * s->ldh));
  NMFB_TRY(ar->alloc(h, &s->pcoef, KTp));
  NMFB_TRY(ar->alloc(h, &s->qcoef, KTp));
  NMFB_TRY(ar->alloc(h, &s->bvec, KTp));
  NMFB_TRY(ar->alloc(h, &s->hscale, KTp));
  NMFB_TRY(ar->alloc(h, &s->ab, 4 * KTp));  // [ab | norm2 | wsum]: WStepArgs::acc
  s->norm2 = s->ab + 2 * KTp;
  s->wsum = s->ab + 3 * KTp;
  if (!s->multi) NMFB_TRY(ar->alloc(h, &s->scal, 8));
  NMFB_TRY(ar->alloc(h, &s->cost, static_cast<size_t>(s->maxiter) + 1));
  NMFB_TRY(ar->alloc(h, &s->stop, 2));
  NMFB_TRY(ar->alloc(h, &s->ticket, 2));
  s->fold_stacks = !s->multi && std::getenv("NMFB_CNMF_HSTACK") == nullptr;

  {  // initial factors: defaults cnmf.m:311, 331-335 (rand; W normalised per basis)
    std::vector<float> tmp;
    const float* Wsrc = cfg.W_init;
    if (!Wsrc) {
      tmp.resize(static_cast<size_t>(m) * KT);
      fill_uniform(tmp, cfg.seed * 2 + 1, false);
      Wsrc = tmp.data();
    }
    NMFB_TRY(upload_colmajor(h, Wsrc, m, KT, s->Wm, s->ldw));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
    const float* Hsrc = cfg.H_init;
    if (!Hsrc) {
      tmp.resize(static_cast<size_t>(K) * n);
      fill_uniform(tmp, cfg.seed * 2 + 2, true);
      Hsrc = tmp.data();
    }
    NMFB_TRY(upload_H(h, ar, Hsrc, K, n, s->Hm, s->ldhm));
    NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  // cnmf.m:157-166: W(:,k,:) /= w_norm, H(k,:) *= w_norm with w_norm = |W(:,k,:)|_F / T
  vec_sums_kernel<<<vec_grid(m, KT), 256, 0, h->stream>>>(s->Wm, KT, m, s->ldw, nullptr, s->norm2, nullptr);
  NMFB_TRY(check_launch(h, "vec_sums(W init)"));
  w_normalize_kernel<<<vec_grid(m, KT), 256, 0, h->stream>>>(s->Wm, s->Wt, m, s->ldw, K, T, 1, s->norm2,
                                                             s->wsum, s->hscale, nullptr);
  NMFB_TRY(check_launch(h, "w_normalize(init)"));
  row_scale_kernel<<<vec_grid(n, K), 256, 0, h->stream>>>(s->Hm, nullptr, n, s->ldhm, s->hscale, 0);
  NMFB_TRY(check_launch(h, "row_scale(H init)"));

  double* sq = nullptr;
  NMFB_TRY(ar->alloc(h, &sq, 1));
  NMFB_TRY(prepare_v_work(h, false, true, sq, nullptr));
  s->Vmma = h->Vwork;
  if (s->multi) {
    NMFB_TRY(comm_allreduce(h, nullptr, 0, sq, 1, nullptr, 0));  // |V|^2 over all column shards
    // V with the T-1 columns that follow the shard (cnmf.m:219 shifts V left): own columns + the right
    // neighbour's first ones, fetched once through the peer region
    NMFB_TRY(ar->alloc(h, &s->Vx, static_cast<size_t>(s->nx) * h->ldv));
    NMFB_CUDA(h, cudaMemcpyAsync(s->Vx, h->Vwork, static_cast<size_t>(n) * h->ldv * sizeof(float), cudaMemcpyDeviceToDevice,
                                 h->stream));
    if (s->halo > 0) {
      NMFB_CUDA(h, cudaMemcpyAsync(region + vhalo_off, h->Vwork, static_cast<size_t>(s->halo) * h->ldv * sizeof(float),
                                   cudaMemcpyDeviceToDevice, h->stream));
      NMFB_TRY(comm_allreduce(h, nullptr, 0, nullptr, 0, s->scal, 4));  // (zeros) - used as a barrier: halos are published
      PeerTable t;
      comm_peer_table(h, &t);
      if (s->hR > 0)
        NMFB_CUDA(h, cudaMemcpyAsync(s->Vx + static_cast<size_t>(n) * h->ldv,
                                     t.base[s->rank + 1] + comm_region_offset(h, region) + vhalo_off,
                                     static_cast<size_t>(s->hR) * h->ldv * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    }
    s->Vmma = s->Vx;
  }
  NMFB_CUDA(h, cudaMemcpyAsync(&s->vsq, sq, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));

  const int* stop = s->stop;
  if (s->two_weight) {
    const bool dual = cfg.alpha == 0;  // cnmf.m:149-153
    s->expo = static_cast<float>(1.0 / (dual ? cfg.beta : cfg.alpha));
    s->ab_scale = -1.0 / (cfg.alpha * cfg.beta);
    NMFB_TRY(ar->alloc(h, &s->Q, static_cast<size_t>(n) * h->ldv));
    NMFB_TRY(ar->alloc(h, &s->Q2, static_cast<size_t>(n) * h->ldv));
    MatRef Xs{s->Wt, m, KTp, s->ldw, true};
    MatRef Ys{s->Hs, n, KTp, s->ldh, true};
    NMFB_TRY(plan_fused(h, &s->gemmS, EPI_ABQ, Xs, Ys, KTp, nullptr, nullptr, 0, m, round_up(n, 64), n, stop));
    GemmArgs& q = s->gemmS.L.args;
    q.Vsrc = h->Vraw;
    q.Qout = s->Q;
    q.Qout2 = s->Q2;
    q.ldv = h->ldv;
    q.scal = s->scal + 2;
    q.ab_mode = cfg.divergence == NMFB_DIV_IS ? ABQ_IS : dual ? ABQ_AB_DUAL : ABQ_AB;
    q.ab_alpha = static_cast<float>(cfg.alpha);
    q.ab_beta = static_cast<float>(cfg.beta);
    q.ab_cost_kl = cfg.divergence == NMFB_DIV_KL ? 1 : 0;
    std::string pe = set_v_prefetch(&s->gemmS.L, h->Vraw, m, n, h->ldv);
    if (!pe.empty()) return h->fail(NMFB_ERR_CUDA, "%s", pe.c_str());
    MatRef Yh{s->Hs, n, KTp, s->ldh, false};
    MatRef Xn{s->Q, m, n, h->ldv, true}, Xp{s->Q2, m, n, h->ldv, true};
    const int tiles = (m + kTileM - 1) / kTileM * ((KTp + kMaxN - 1) / kMaxN);
    const bool split = tiles * 2 <= h->num_sms;
    NMFB_TRY(plan_store(h, ar, &s->gemmRa, Xn, Yh, n, nullptr, nullptr, 0, m, KTp, s->A, nullptr, s->ldw, split, stop));
    NMFB_TRY(plan_store(h, ar, &s->gemmRb, Xp, Yh, n, nullptr, nullptr, 0, m, KTp, s->B, nullptr, s->ldw, split, stop));
    MatRef Yw{s->Wt, m, KTp, s->ldw, false};
    MatRef Xnt{s->Q, m, n, h->ldv, false}, Xpt{s->Q2, m, n, h->ldv, false};
    const int tiles_h = (n + kTileM - 1) / kTileM * ((KTp + kMaxN - 1) / kMaxN);
    const bool split_h = tiles_h * 2 <= h->num_sms;
    NMFB_TRY(plan_store(h, ar, &s->gemmPn, Xnt, Yw, m, nullptr, nullptr, 0, n, KTp, s->P, nullptr, s->ldh, split_h, stop));
    NMFB_TRY(plan_store(h, ar, &s->gemmPd, Xpt, Yw, m, nullptr, nullptr, 0, n, KTp, s->D, nullptr, s->ldh, split_h, stop));
    loop_begin(h);
    NMFB_TRY(run_chunked(h, s->maxiter, s->stop, [&](int i) { return enqueue_iteration(h, s, i); }));
    loop_end(h, s->maxiter);
    NMFB_TRY(enqueue_hstack(h, s, stop));
    s->gemmS.L.args.want_cost = 1;
    NMFB_TRY(run_gemm(h, s->gemmS));
    NMFB_TRY(enqueue_cost(h, s, s->maxiter - 1));
    return cnmf_finish(h, s, W_out, H_out, cost_out, n_cost);
  }
  NMFB_TRY(plan_gram(h, ar, &s->gramW, s->Wt, KTp, m, s->ldw, stop));
  NMFB_TRY(plan_gram(h, ar, &s->gramH, s->Hs, KTp, n, s->ldh, stop));  // own columns only
  if (s->multi) s->gramH.g32 = s->packed + static_cast<size_t>(KTp) * s->ldw;  // behind A: one all-reduce
  {
    const int nx = s->nx;  // own + right-halo columns
    MatRef Xv{s->Vmma, m, n, h->ldv, true};
    MatRef Yh{s->Hs, n, KTp, s->ldh, false};
    MatRef Xw{s->Wt, m, KTp, s->ldw, true};
    MatRef Yg{s->gramH.gtf, KTp, KTp, KTp, false};
    const int tiles = (m + kTileM - 1) / kTileM * ((KTp + kMaxN - 1) / kMaxN);
    if (s->multi) {  // A is a partial sum: B = Wc (Hs Hs') waits for the all-reduce
      NMFB_TRY(plan_store(h, ar, &s->gemmA, Xv, Yh, n, nullptr, nullptr, 0, m, KTp, s->A, nullptr, s->ldw,
                          tiles * 2 <= h->num_sms, stop));
      NMFB_TRY(plan_store(h, ar, &s->gemmB, Xw, Yg, KTp, nullptr, nullptr, 0, m, KTp, s->B, nullptr, s->ldw, false, stop));
    } else {
      NMFB_TRY(plan_store(h, ar, &s->gemmA, Xv, Yh, n, &Xw, &Yg, KTp, m, KTp, s->A, s->B, s->ldw,
                          tiles * 2 <= h->num_sms, stop));
    }
    MatRef Xvt{s->Vmma, m, nx, h->ldv, false};
    MatRef Yw{s->Wt, m, KTp, s->ldw, false};
    MatRef Xh{s->Hs, nx, KTp, s->ldh, true};
    MatRef Ygw{s->gramW.gtf, KTp, KTp, KTp, false};
    const int tiles_h = (nx + kTileM - 1) / kTileM * ((KTp + kMaxN - 1) / kMaxN);
    int p_tile_n = 0;  // experiment: narrower tiles against the 2.1-wave quantisation of the P, D contraction
    if (const char* e = std::getenv("NMFB_CNMF_TILEN")) p_tile_n = std::atoi(e);
    NMFB_TRY(plan_store(h, ar, &s->gemmP, Xvt, Yw, m, &Xh, &Ygw, KTp, nx, KTp, s->P, s->D, s->ldh,
                        tiles_h * 2 <= h->num_sms, stop, nullptr, p_tile_n));
  }
  if (s->W_fixed) NMFB_TRY(run_gram(h, s->gramW, nullptr));
  // sum(H) for the sparsity term when H is never updated is accumulated by fold_update (freeze)

  loop_begin(h);
  NMFB_TRY(run_chunked(h, s->maxiter, s->stop, [&](int i) { return enqueue_iteration(h, s, i); }));
  loop_end(h, s->maxiter);
  // cost of the last executed iteration needs Hs Hs' of the final H
  NMFB_TRY(enqueue_hstack(h, s, stop));
  NMFB_TRY(run_gram(h, s->gramH, stop));
  if (s->multi)
    NMFB_TRY(comm_allreduce(h, s->gramH.g32, static_cast<size_t>(KTp) * KTp, nullptr, 0, s->scal, 4));
  NMFB_TRY(enqueue_cost(h, s, s->maxiter - 1));
  return cnmf_finish(h, s, W_out, H_out, cost_out, n_cost);
}

}  // namespace cnmfdetail

extern "C" int nmfb_cnmf(nmfb_handle* h, int K, int T, const nmfb_config* cfg, float* W_out, float* H_out,
                         double* cost_out, int* n_cost) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  cnmfdetail::CnmfState st;
  int rc = cnmfdetail::cnmf_run(h, &st, K, T, cfg, W_out, H_out, cost_out, n_cost);
  cudaStreamSynchronize(h->stream);
  return rc;
}
