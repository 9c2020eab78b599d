// Host side of kl_fused_kernel: tensor maps, launch geometry, the two small finishing kernels.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "engine.cuh"
#include "ew_kernels.cuh"
#include "kl_fused.cuh"
#include "ab_fused.cuh"
#include "resid_fused.cuh"

namespace nmfb {

struct KlOp {
  CUtensorMap tmF, tmG1, tmG2, tmV;
  KlArgs args;
  dim3 grid;
  int splits = 1;
  float* parts = nullptr;
  bool planned = false;
};

// F: [Kp][ldf] rows contiguous (length rows); G: [Kp][ldg] (length cols); VT: [cols][ldvt] (rows contiguous)
int plan_kl(nmfb_handle* h, Arena* ar, KlOp* op, const float* F, long long ldf, const float* G, long long ldg,
            const float* VT, long long ldvt, int rows, int cols, int Kp, const int* stop) {
  if (Kp % 32 != 0 || Kp > kKlMaxKp) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "plan_kl: Kp must be 32..128");
  std::string e;
  if (!(e = make_tmap(&op->tmF, Mat2D{F, rows, Kp, ldf}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "kl F %s", e.c_str());
  if (!(e = make_tmap(&op->tmG1, Mat2D{G, cols, Kp, ldg}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "kl G1 %s", e.c_str());
  if (!(e = make_tmap(&op->tmG2, Mat2D{G, cols, Kp, ldg}, 32, Kp / 2, false)).empty()) return h->fail(NMFB_ERR_CUDA, "kl G2 %s", e.c_str());
  if (!(e = make_tmap(&op->tmV, Mat2D{VT, rows, cols, ldvt}, kTileM, kKlTileC, false, true)).empty())
    return h->fail(NMFB_ERR_CUDA, "kl V %s", e.c_str());
  const int pairs = (rows + 2 * kTileM - 1) / (2 * kTileM);
  const int total_tiles = (cols + kKlTileC - 1) / kKlTileC;
  int per = 0;
  const int splits = choose_kl_splits(pairs, total_tiles, rows, Kp, h->num_sms / 2, kKlOutChunk, &per);
  op->splits = splits;
  op->grid = dim3(2 * pairs, splits, 1);
  KlArgs& a = op->args;
  a.rows = rows;
  a.cols = cols;
  a.Kp = Kp;
  a.tiles_per_split = per;
  a.want_cost = 0;
  a.ldo = (rows + 3) / 4 * 4;
  a.slab = static_cast<long long>(Kp) * a.ldo;
  a.scal = nullptr;
  a.stop = stop;
  NMFB_TRY(ar->alloc(h, &op->parts, static_cast<size_t>(splits) * a.slab));
  a.out = op->parts;
  op->planned = true;
  return NMFB_OK;
}

int run_kl(nmfb_handle* h, const KlOp& op) {
  static cudaError_t attr = cudaFuncSetAttribute(kl_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKlSmemBytes);
  if (attr != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "cudaFuncSetAttribute(kl_fused): %s", cudaGetErrorString(attr));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = op.grid;
  cfg.blockDim = dim3(kKlThreads);
  cfg.dynamicSmemBytes = kKlSmemBytes;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kl_fused_kernel, op.tmF, op.tmG1, op.tmG2, op.tmV, op.args);
  ++h->launches;
  if (e != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "launch of kl_fused failed: %s", cudaGetErrorString(e));
  return NMFB_OK;
}

// ---------------------------------------------------------------- ab_fused (IS / AB divergences)
struct AbOp {
  CUtensorMap tmF, tmG1, tmG2, tmV;
  AbArgs args;
  dim3 grid;
  int splits = 1;
  int mode = ABQ_IS;
  float* parts_n = nullptr;
  float* parts_p = nullptr;
  bool planned = false;
};

// Same operand conventions as plan_kl; mode = ABQ_IS / ABQ_AB / ABQ_AB_DUAL.
int plan_ab(nmfb_handle* h, Arena* ar, AbOp* op, const float* F, long long ldf, const float* G, long long ldg,
            const float* VT, long long ldvt, int rows, int cols, int Kp, int mode, float alpha, float beta,
            const int* stop) {
  if (Kp % 32 != 0 || Kp > kKlMaxKp) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "plan_ab: Kp must be 32..128");
  std::string e;
  if (!(e = make_tmap(&op->tmF, Mat2D{F, rows, Kp, ldf}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "ab F %s", e.c_str());
  if (!(e = make_tmap(&op->tmG1, Mat2D{G, cols, Kp, ldg}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "ab G1 %s", e.c_str());
  if (!(e = make_tmap(&op->tmG2, Mat2D{G, cols, Kp, ldg}, 32, Kp / 2, false)).empty()) return h->fail(NMFB_ERR_CUDA, "ab G2 %s", e.c_str());
  if (!(e = make_tmap(&op->tmV, Mat2D{VT, rows, cols, ldvt}, kTileM, kKlTileC, false, true)).empty())
    return h->fail(NMFB_ERR_CUDA, "ab V %s", e.c_str());
  const int pairs = (rows + 2 * kTileM - 1) / (2 * kTileM);
  const int total_tiles = (cols + kKlTileC - 1) / kKlTileC;
  int max_tiles = kAbMaxTiles;
  if (const char* env = std::getenv("NMFB_AB_MAX_TILES")) max_tiles = std::max(kKlOutChunk, std::atoi(env));  // experiments
  int per = 0;
  const int splits = choose_kl_splits(pairs, total_tiles, rows, 2 * Kp, h->num_sms / 2, kKlOutChunk, &per, max_tiles);
  op->splits = splits;
  op->mode = mode;
  op->grid = dim3(2 * pairs, splits, 1);
  AbArgs& a = op->args;
  a.rows = rows;
  a.cols = cols;
  a.Kp = Kp;
  a.tiles_per_split = per;
  a.want_cost = 0;
  a.alpha = alpha;
  a.beta = beta;
  a.ldo = (rows + 3) / 4 * 4;
  a.slab = static_cast<long long>(Kp) * a.ldo;
  a.scal = nullptr;
  a.stop = stop;
  NMFB_TRY(ar->alloc(h, &op->parts_n, static_cast<size_t>(splits) * a.slab));
  NMFB_TRY(ar->alloc(h, &op->parts_p, static_cast<size_t>(splits) * a.slab));
  a.out_n = op->parts_n;
  a.out_p = op->parts_p;
  op->planned = true;
  return NMFB_OK;
}

int run_ab(nmfb_handle* h, const AbOp& op) {
  static cudaError_t attr = []() {
    cudaError_t e0 = cudaFuncSetAttribute(ab_fused_kernel<ABQ_IS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes);
    cudaError_t e1 = cudaFuncSetAttribute(ab_fused_kernel<ABQ_AB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes);
    cudaError_t e2 = cudaFuncSetAttribute(ab_fused_kernel<ABQ_AB_DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes);
    return e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2);
  }();
  if (attr != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "cudaFuncSetAttribute(ab_fused): %s", cudaGetErrorString(attr));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = op.grid;
  cfg.blockDim = dim3(kKlThreads);
  cfg.dynamicSmemBytes = kAbSmemBytes;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaSuccess;
  switch (op.mode) {
    case ABQ_IS: e = cudaLaunchKernelEx(&cfg, ab_fused_kernel<ABQ_IS>, op.tmF, op.tmG1, op.tmG2, op.tmV, op.args); break;
    case ABQ_AB: e = cudaLaunchKernelEx(&cfg, ab_fused_kernel<ABQ_AB>, op.tmF, op.tmG1, op.tmG2, op.tmV, op.args); break;
    default: e = cudaLaunchKernelEx(&cfg, ab_fused_kernel<ABQ_AB_DUAL>, op.tmF, op.tmG1, op.tmG2, op.tmV, op.args); break;
  }
  ++h->launches;
  if (e != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "launch of ab_fused failed: %s", cudaGetErrorString(e));
  return NMFB_OK;
}

// ---------------------------------------------------------------- resid_fused (nmfsc objective)
struct ResidOp {
  CUtensorMap tmFhi, tmFlo, tmGhi, tmGlo, tmV;
  ResidArgs args;
  dim3 grid;
  bool planned = false;
};

// W (head, tail): [Kp][ldw] columns contiguous over m rows; H (head, tail): [Kp][ldh] over n columns;
// V: column-major m x n (leading dimension ldv); scal[0] receives sum (V - W H)^2.
int plan_resid(nmfb_handle* h, ResidOp* op, const float* Whi, const float* Wlo, long long ldw, const float* Hhi,
               const float* Hlo, long long ldh, const float* V, long long ldv, int m, int n, int Kp, double* scal) {
  if (Kp % 32 != 0 || Kp > kKlMaxKp) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "plan_resid: Kp must be 32..128");
  std::string e;
  if (!(e = make_tmap(&op->tmFhi, Mat2D{Whi, m, Kp, ldw}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "resid W %s", e.c_str());
  if (!(e = make_tmap(&op->tmFlo, Mat2D{Wlo, m, Kp, ldw}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "resid W %s", e.c_str());
  if (!(e = make_tmap(&op->tmGhi, Mat2D{Hhi, n, Kp, ldh}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "resid H %s", e.c_str());
  if (!(e = make_tmap(&op->tmGlo, Mat2D{Hlo, n, Kp, ldh}, 32, 32, true)).empty()) return h->fail(NMFB_ERR_CUDA, "resid H %s", e.c_str());
  if (!(e = make_tmap(&op->tmV, Mat2D{V, m, n, ldv}, kTileM, kKlTileC, false, true)).empty())
    return h->fail(NMFB_ERR_CUDA, "resid V %s", e.c_str());
  const int pairs = (m + 2 * kTileM - 1) / (2 * kTileM);
  const int total_tiles = (n + kKlTileC - 1) / kKlTileC;
  int splits = std::max(1, (h->num_sms / 2) / pairs);
  splits = std::min(splits, total_tiles);
  const int per = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + per - 1) / per;
  op->grid = dim3(2 * pairs, splits, 1);
  op->args = ResidArgs{m, n, Kp, per, scal, nullptr, LsFin()};
  op->planned = true;
  return NMFB_OK;
}

int run_resid(nmfb_handle* h, const ResidOp& op) {
  static cudaError_t attr =
      cudaFuncSetAttribute(resid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsSmemBytes);
  if (attr != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "cudaFuncSetAttribute(resid_fused): %s", cudaGetErrorString(attr));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = op.grid;
  cfg.blockDim = dim3(64 + kKlEpiWarps * 32);
  cfg.dynamicSmemBytes = kRsSmemBytes;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, resid_fused_kernel, op.tmFhi, op.tmFlo, op.tmGhi, op.tmGlo, op.tmV, op.args);
  ++h->launches;
  if (e != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "launch of resid_fused failed: %s", cudaGetErrorString(e));
  return NMFB_OK;
}

// H half: N = sum of the column-split slabs; H <- H .* N ./ max(ws + lambda, eps)  (nmf.m:183-184,199)
__global__ void kl_h_finish_kernel(const float* __restrict__ parts, int splits, long long slab, long long ldo,
                                   float* __restrict__ Hm, float* __restrict__ Ht, long long ldh,
                                   const float* __restrict__ ws, float lambda, int n, int freeze, double* scal,
                                   const int* stop, const float* lambda_k = nullptr, const int* fixed_k = nullptr,
                                   int lnmf = 0) {
  NMFB_STOP_GUARD(stop);
  __shared__ double sh[32];
  const int k = blockIdx.y;
  if (lambda_k != nullptr) lambda = lambda_k[k];  // multi-source run: scal[1] receives the weighted sum
  if (fixed_k != nullptr && fixed_k[k] != 0) freeze = 1;
  const double wgt = lambda_k != nullptr ? static_cast<double>(lambda) : 1.0;
  const float den = fmaxf(ws[k] + lambda, NMFB_EPS);
  double acc[1] = {0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    float nv = 0.f;
    for (int z = 0; z < splits; ++z) nv += parts[z * slab + k * ldo + j];
    float hv = Hm[k * ldh + j];
    if (!freeze) {
      hv = lnmf ? sqrtf(hv * nv) : hv * (nv / den);  // lnmf.m:81 / nmf.m:183-184,199
      Hm[k * ldh + j] = hv;
      Ht[k * ldh + j] = tf32_rn(hv);
    }
    acc[0] += hv;
  }
  acc[0] *= wgt;
  block_sum<1>(acc, sh);
  if (threadIdx.x == 0) atomicAdd(scal + 1, acc[0]);
}

}  // namespace nmfb
