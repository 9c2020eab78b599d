// Kernel-level test hooks (C ABI, device pointers in/out).  Used only by tests/ to check panel_gemm against a
// plain matmul on the same GPU.  NOT part of libnmfb200.so: built into tests/libnmfb200_test.so together with
// gemm_host.cu (csrc/testlib.cu), so the product library exports nothing but include/nmfb200.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <vector>

#include "gemm_host.cuh"

using namespace nmfb;

// NMFB_DEBUG_CG=2 runs the kernel-level tests on the CTA-pair kernel, 4 on clusters of two pairs sharing the Y
// slab by TMA multicast (shapes that kernel cannot take - odd number of 256-row tiles, slabs that are not
// 128 / 256 columns wide - fall back to plain pairs)
static int debug_cg(int rows = 0, int ncols = 0) {
  const char* e = std::getenv("NMFB_DEBUG_CG");
  int cg = (e && e[0] == '4') ? 4 : (e && e[0] == '2') ? 2 : 1;
  if (cg == 4) {
    const int tiles = (rows + 2 * kTileM - 1) / (2 * kTileM);
    if (tiles % 2 != 0 || std::min(ncols, kMaxN) % 128 != 0) cg = 2;
  }
  return cg;
}

// NMFB_DEBUG_TAIL=<helpers>: the pair kernel runs with that many tail-helper pairs (GemmArgs::sk_*) whenever the
// launch has one column chunk, no split-K and at least two k-blocks; the scratch lives until `release`
struct DebugTail {
  float* part = nullptr;
  unsigned int* flags = nullptr;
  void release() {
    cudaFree(part);
    cudaFree(flags);
    part = nullptr;
    flags = nullptr;
  }
};
static void debug_tail(GemmLaunch* L, DebugTail* t) {
  const char* e = std::getenv("NMFB_DEBUG_TAIL");
  if (!e) return;
  const GemmArgs& a = L->args;
  int helpers = std::atoi(e);
  if (helpers <= 0 || L->cg != 2 || L->grid.y != 1 || L->grid.z != 1 || a.nkb0 < 2 || a.nkb_seg != a.nkb0) return;
  const int tiles = static_cast<int>(L->grid.x) / 2;
  helpers = std::min(helpers, tiles);
  const int per = (tiles + helpers - 1) / helpers;
  int kp = (a.nkb0 * per + per) / (per + 1);
  kp = std::min(std::max(kp, 1), a.nkb0 - 1);
  const size_t part_bytes = static_cast<size_t>(tiles) * a.ncols * 2 * kTileM * sizeof(float);
  if (cudaMalloc(&t->part, part_bytes) != cudaSuccess) return;
  if (cudaMalloc(&t->flags, 2 * tiles * sizeof(unsigned int)) != cudaSuccess) return;
  cudaMemset(t->part, 0xff, part_bytes);  // NaN: a primary that does not wait for its helper is caught
  cudaMemset(t->flags, 0, 2 * tiles * sizeof(unsigned int));
  set_tail_helpers(L, helpers, kp, t->part, t->flags);
  L->args.sk_epoch = 1;
}

static int fail(char* err, int errlen, const std::string& msg) {
  if (err && errlen > 0) {
    std::snprintf(err, errlen, "%s", msg.c_str());
  }
  return 1;
}

extern "C" {

struct nmfb_debug_mat {
  const float* base;
  long long inner, outer, pitch;
  int mn_major;
};

// out0[z*split_stride + col*ldo + row] = sum_k X0[row,k] * Y0[col,k]  (per split z)
// out1[col*ldo + row]                  = sum_k X1[row,k] * Y1[col,k]  (if X1.base != NULL)
int nmfb_debug_gemm_store(const nmfb_debug_mat* X0, const nmfb_debug_mat* Y0, long long kdim0,
                          const nmfb_debug_mat* X1, const nmfb_debug_mat* Y1, long long kdim1,
                          int rows, int ncols, int splits, float* out0, float* out1, long long ldo,
                          long long split_stride, int* splits_used, char* err, int errlen) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GemmLaunch L;
  GemmOperand x0{{X0->base, X0->inner, X0->outer, X0->pitch}, X0->mn_major != 0};
  GemmOperand y0{{Y0->base, Y0->inner, Y0->outer, Y0->pitch}, Y0->mn_major != 0};
  GemmOperand x1{};
  GemmOperand y1{};
  const bool two = X1 && X1->base;
  if (two) {
    x1 = GemmOperand{{X1->base, X1->inner, X1->outer, X1->pitch}, X1->mn_major != 0};
    y1 = GemmOperand{{Y1->base, Y1->inner, Y1->outer, Y1->pitch}, Y1->mn_major != 0};
  }
  std::string e = plan_gemm(&L, x0, y0, kdim0, two ? &x1 : nullptr, two ? &y1 : nullptr, kdim1, rows,
                            ncols, splits, sms, debug_cg(rows, ncols));
  if (!e.empty()) return fail(err, errlen, e);
  L.args.out0 = out0;
  L.args.out1 = out1;
  L.args.ldo = ldo;
  L.args.split_stride = split_stride;
  if (splits_used) *splits_used = static_cast<int>(L.grid.z);
  DebugTail tail;
  debug_tail(&L, &tail);
  e = launch_gemm(L, EPI_STORE, 0);
  if (!e.empty()) return fail(err, errlen, e);
  cudaError_t ce = cudaDeviceSynchronize();
  tail.release();
  if (ce != cudaSuccess) return fail(err, errlen, std::string("sync: ") + cudaGetErrorString(ce));
  return 0;
}

// Fused H update: acc0 = X0*Y0^T (N), acc1 = X1*Y1^T (D); H updated in place.
int nmfb_debug_gemm_hupdate(const nmfb_debug_mat* X0, const nmfb_debug_mat* Y0, long long kdim0,
                            const nmfb_debug_mat* X1, const nmfb_debug_mat* Y1, long long kdim1,
                            int rows, int ncols, float* Hm, float* Hr32, float* Hc32, long long ldh,
                            long long ldc, float lambda, double* partials, char* err, int errlen) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GemmLaunch L;
  GemmOperand x0{{X0->base, X0->inner, X0->outer, X0->pitch}, X0->mn_major != 0};
  GemmOperand y0{{Y0->base, Y0->inner, Y0->outer, Y0->pitch}, Y0->mn_major != 0};
  GemmOperand x1{{X1->base, X1->inner, X1->outer, X1->pitch}, X1->mn_major != 0};
  GemmOperand y1{{Y1->base, Y1->inner, Y1->outer, Y1->pitch}, Y1->mn_major != 0};
  std::string e = plan_gemm(&L, x0, y0, kdim0, &x1, &y1, kdim1, rows, ncols, 1, sms, debug_cg(rows, ncols));
  if (!e.empty()) return fail(err, errlen, e);
  L.args.Hm = Hm;
  L.args.Hr32 = Hr32;
  L.args.Hc32 = Hc32;
  L.args.ldh = ldh;
  L.args.ldc = ldc;
  L.args.lambda = lambda;
  L.args.scal = partials;
  DebugTail tail;
  debug_tail(&L, &tail);
  e = launch_gemm(L, EPI_HUPDATE, 0);
  if (!e.empty()) return fail(err, errlen, e);
  cudaError_t ce = cudaDeviceSynchronize();
  tail.release();
  if (ce != cudaSuccess) return fail(err, errlen, std::string("sync: ") + cudaGetErrorString(ce));
  return 0;
}

// Host-side planners (pure functions): exposed so that the CPU tests can check them without a GPU.
int nmfb_debug_kl_splits(int pairs, int total_tiles, int rows, int Kp, int slots, int chunk, int max_per, int* per_out) {
  return choose_kl_splits(pairs, total_tiles, rows, Kp, slots, chunk, per_out, max_per);
}

int nmfb_debug_tail_balance(int tiles, int helpers, int nkb0, int* kp_out) {
  return balance_tail_helpers(tiles, helpers, nkb0, kp_out);
}

}  // extern "C"
