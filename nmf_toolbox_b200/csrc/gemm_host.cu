#include "gemm_host.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace nmfb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  static std::string saved_err;
  std::call_once(once, [&]() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      saved_err = std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") +
                  cudaGetErrorString(e);
    } else {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  if (!fn && err) *err = saved_err;
  return fn;
}

std::string make_tmap(CUtensorMap* out, const Mat2D& m, uint32_t box_inner, uint32_t box_outer,
                      bool atom32b, bool no_swizzle) {
  std::string err;
  EncodeTiledFn fn = get_encode_fn(&err);
  if (!fn) return err;
  if ((m.pitch * 4) % 16 != 0) return "tensor map: row pitch is not a multiple of 16 bytes";
  if ((reinterpret_cast<uintptr_t>(m.base) & 15) != 0) return "tensor map: base not 16-byte aligned";
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(m.inner), static_cast<cuuint64_t>(m.outer)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(m.pitch) * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(m.base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                             : (atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
           " (inner=" + std::to_string(m.inner) + " outer=" + std::to_string(m.outer) +
           " pitch=" + std::to_string(m.pitch) + " box=" + std::to_string(box_inner) + "x" +
           std::to_string(box_outer) + ")";
  }
  return "";
}

int choose_splits(int tiles, int nkb0, int num_sms, int* kb_per_split) {
  int splits = std::max(1, num_sms / std::max(1, tiles));
  splits = std::min(splits, std::max(1, nkb0));
  int per = (nkb0 + splits - 1) / std::max(1, splits);
  per = std::max(per, 1);
  splits = std::max(1, (nkb0 + per - 1) / per);
  *kb_per_split = per;
  return splits;
}

bool pair_eligible(int rows, int ncols, bool y_mn_major0, bool y_mn_major1) {
  const char* env = std::getenv("NMFB_CTA_GROUP");
  if (env && env[0] == '1') return false;
  if (rows <= kTileM) return false;  // a single 128-row tile: the partner CTA would idle
  if ((y_mn_major0 || y_mn_major1) && ncols % 64 != 0) return false;  // 32-row boxes per CTA half
  return true;
}

int choose_cg(int epi, int rows, int ncols, bool y_mn_major0, bool y_mn_major1, int tile_n) {
  if (!pair_eligible(rows, ncols, y_mn_major0, y_mn_major1)) return 1;
  // Two pairs sharing the Y slab by TMA multicast (cluster of four): the Euclidean iteration's two kernels only,
  // an even number of 256-row tiles, slabs of 128 / 256 columns.  NMFB_CTA_GROUP=2 keeps plain pairs.
  const char* env = std::getenv("NMFB_CTA_GROUP");
  const bool want4 = env ? env[0] == '4' : false;
  const int tiles = (rows + 2 * kTileM - 1) / (2 * kTileM);
  const int box_n = std::min(ncols, tile_n > 0 ? tile_n : kMaxN);
  if (want4 && (epi == EPI_STORE || epi == EPI_HUPDATE) && tiles % 2 == 0 && box_n % 128 == 0) return 4;
  return 2;
}

std::string plan_gemm(GemmLaunch* L, const GemmOperand& X0, const GemmOperand& Y0, long long kdim0,
                      const GemmOperand* X1, const GemmOperand* Y1, long long kdim1, int rows,
                      int ncols, int splits_hint, int num_sms, int cg, int tile_n) {
  if (ncols <= 0 || ncols % 32 != 0) return "plan_gemm: ncols must be a positive multiple of 32";
  if (cg != 1 && cg != 2 && cg != 4) return "plan_gemm: cta group must be 1, 2 or 4 (two pairs sharing the Y slab)";
  const int pair = cg >= 2 ? 2 : 1;
  std::memset(L, 0, sizeof(*L));
  L->cg = cg;
  GemmArgs& a = L->args;
  a.rows = rows;
  a.ncols = ncols;
  if (tile_n < 0 || tile_n > kMaxN || tile_n % 32 != 0) return "plan_gemm: tile_n must be a multiple of 32 up to 256";
  a.tile_n = tile_n;
  const int tn = tile_n > 0 ? tile_n : kMaxN;
  a.box_n = std::min(ncols, tn);
  a.nkb0 = static_cast<int>((kdim0 + kBlockK - 1) / kBlockK);
  a.nkb_seg = a.nkb0;
  a.chunk_kb = 0;
  if (const char* env = std::getenv("NMFB_CHUNK_KB")) a.chunk_kb = std::atoi(env);  // tuning experiments
  a.nkb1 = (X1 != nullptr) ? static_cast<int>((kdim1 + kBlockK - 1) / kBlockK) : 0;
  a.xmn0 = X0.mn_major ? 1 : 0;
  a.xmn1 = (X1 && X1->mn_major) ? 1 : 0;
  a.ymn0 = Y0.mn_major ? 1 : 0;
  a.ymn1 = (Y1 && Y1->mn_major) ? 1 : 0;
  a.ncols_valid = ncols;
  const int tile_rows = kTileM * pair;
  const int tiles = (rows + tile_rows - 1) / tile_rows;
  if (cg == 4 && (tiles % 2 != 0 || a.box_n % 128 != 0))
    return "plan_gemm: the cluster-of-four kernel needs an even number of 256-row tiles and slabs of 128 or 256 columns";
  const int chunks = (ncols + tn - 1) / tn;
  int splits = 1;
  a.kb_per_split = std::max(a.nkb0, 1);
  if (splits_hint != 1) {
    if (splits_hint <= 0) {
      splits = choose_splits(tiles * chunks * pair, a.nkb0, num_sms, &a.kb_per_split);
    } else {
      int per = std::max(1, (a.nkb0 + splits_hint - 1) / splits_hint);
      splits = std::max(1, (a.nkb0 + per - 1) / per);
      a.kb_per_split = per;
    }
  }
  L->grid = dim3(tiles * pair, chunks, splits);

  std::string e;
  auto xmap = [&](CUtensorMap* tm, const GemmOperand& X) {
    return X.mn_major ? make_tmap(tm, X.m, 32, 32, true) : make_tmap(tm, X.m, kBlockK, kTileM, false);
  };
  if (!(e = xmap(&L->tmX0, X0)).empty()) return "X0 " + e;
  auto ymap = [&](CUtensorMap* tm, const GemmOperand& Y) {  // each CTA of a pair loads half of the slab
    return Y.mn_major ? make_tmap(tm, Y.m, 32, 32, true)
                      : make_tmap(tm, Y.m, kBlockK, a.box_n / cg, false);
  };
  if (!(e = ymap(&L->tmY0, Y0)).empty()) return "Y0 " + e;
  if (X1) {
    if (!(e = xmap(&L->tmX1, *X1)).empty()) return "X1 " + e;
    if (!(e = ymap(&L->tmY1, *Y1)).empty()) return "Y1 " + e;
  } else {
    L->tmX1 = L->tmX0;
    L->tmY1 = L->tmY0;
  }
  L->tmXb = L->tmXc = L->tmX0;
  L->tmYb = L->tmYc = L->tmY0;
  L->tmH = L->tmX0;
  return "";
}

std::string set_h_prefetch(GemmLaunch* L, const float* Hm, long long n, long long Kp, long long ldh) {
  Mat2D m{Hm, n, Kp, ldh};
  std::string e = make_tmap(&L->tmH, m, kTileM, 32, false, true);  // 128 samples x 32 basis rows, linear
  if (!e.empty()) return "H tile " + e;
  L->args.h_prefetch = 1;
  return "";
}

std::string set_v_prefetch(GemmLaunch* L, const float* V, long long m, long long n, long long ldv) {
  Mat2D mm{V, m, n, ldv};
  std::string e = make_tmap(&L->tmH, mm, kTileM, 32, false, true);  // 128 rows x 32 columns, linear
  if (!e.empty()) return "V tile " + e;
  L->args.h_prefetch = 1;
  return "";
}

std::string add_segment(GemmLaunch* L, int seg, const GemmOperand& X, const GemmOperand& Y, int num_sms,
                        int splits_hint) {
  GemmArgs& a = L->args;
  if (seg < 1 || seg > 2) return "add_segment: segment index must be 1 or 2";
  if ((X.mn_major ? 1 : 0) != a.xmn0 || (Y.mn_major ? 1 : 0) != a.ymn0)
    return "add_segment: operands must have the layout of segment 0";
  std::string e;
  CUtensorMap* tx = seg == 1 ? &L->tmXb : &L->tmXc;
  CUtensorMap* ty = seg == 1 ? &L->tmYb : &L->tmYc;
  e = X.mn_major ? make_tmap(tx, X.m, 32, 32, true) : make_tmap(tx, X.m, kBlockK, kTileM, false);
  if (!e.empty()) return "Xseg " + e;
  e = Y.mn_major ? make_tmap(ty, Y.m, 32, 32, true) : make_tmap(ty, Y.m, kBlockK, a.box_n / L->cg, false);
  if (!e.empty()) return "Yseg " + e;
  a.nkb0 = a.nkb_seg * (seg + 1);
  // redo the split bookkeeping for the longer contraction
  int splits = 1;
  a.kb_per_split = a.nkb0;
  if (splits_hint != 1) {
    if (splits_hint <= 0) {
      splits = choose_splits(static_cast<int>(L->grid.x * L->grid.y), a.nkb0, num_sms, &a.kb_per_split);
    } else {
      int per = std::max(1, (a.nkb0 + splits_hint - 1) / splits_hint);
      splits = std::max(1, (a.nkb0 + per - 1) / per);
      a.kb_per_split = per;
    }
  }
  L->grid.z = splits;
  return "";
}

// Column splits of a kl_fused / ab_fused launch.  One CTA pair is resident per two SMs (shared memory), so the
// pairs * splits clusters of a launch run as work items on `slots` pair slots, handed out in launch order (all
// row blocks of split 0, then split 1, ...).  32 row blocks x 2 equal splits keep 64 of the 74 slots busy for
// the whole launch; 32 x 3 with splits of 444 + 444 + 136 tiles put the 64 long items on 64 slots and the 32
// short ones on the other 10 (three or four each): same work, every slot busy.  The planner tries every split
// length (a multiple of the accumulation chunk; the last split takes the remainder), plays the launch through
// a list scheduler and keeps the shortest.  Every work item is charged a fixed overhead (launch, TMEM
// allocation, the F tile, pipeline fill, the partial slab, teardown: ~10 column tiles, fitted to measured launches
// with 1, 2 and 9 equal splits) and every split the traffic of one more partial slab.
int choose_kl_splits(int pairs, int total_tiles, int rows, int Kp, int slots, int chunk, int* per_out, int max_per) {
  const char* env = std::getenv("NMFB_KL_SPLITS");  // "0": the round-1 rule (one wave of equal splits); n > 0: n equal splits
  const int forced = env ? std::atoi(env) : -1;
  const double tile_us = 1.2 * Kp / 128.0;                                // measured: 0.62 ms for 512 tiles, Kp = 128
  const double slab_us = static_cast<double>(rows) * Kp * 8.0 / 5.0e6;   // write + read of one slab at ~5 TB/s
  const double item_overhead = 10.0;                                      // in tiles
  auto round_chunk = [chunk](int p) { return (p + chunk - 1) / chunk * chunk; };
  if (max_per <= 0) max_per = total_tiles;
  max_per = std::max(chunk, max_per / chunk * chunk);
  if (forced >= 0) {
    const int s = forced > 0 ? forced : std::max(1, std::min(slots / std::max(1, pairs), total_tiles));
    const int per = std::min(max_per, round_chunk((total_tiles + s - 1) / s));
    *per_out = per;
    return (total_tiles + per - 1) / per;
  }
  int best_per = std::min(max_per, round_chunk(total_tiles));
  double best = 1e300;
  std::vector<double> slot_free(static_cast<size_t>(std::max(1, slots)));
  for (int per = chunk; per <= std::min(max_per, round_chunk(total_tiles)); per += chunk) {
    const int splits = (total_tiles + per - 1) / per;
    if (static_cast<long long>(splits) * pairs > 65535 || splits > 256) continue;
    std::fill(slot_free.begin(), slot_free.end(), 0.0);
    // list scheduling in launch order; the slots form a heap keyed by the time they become free
    auto cmp = [](double x, double y) { return x > y; };
    double makespan = 0.0;
    for (int y = 0; y < splits; ++y) {
      const double len = std::min(per, total_tiles - y * per) + item_overhead;
      for (int p = 0; p < pairs; ++p) {
        std::pop_heap(slot_free.begin(), slot_free.end(), cmp);
        const double done = slot_free.back() + len;
        slot_free.back() = done;
        std::push_heap(slot_free.begin(), slot_free.end(), cmp);
        makespan = std::max(makespan, done);
      }
    }
    const double t = makespan * tile_us + splits * slab_us;
    if (t < best * (1.0 - 1e-9)) {
      best = t;
      best_per = per;
    }
  }
  *per_out = best_per;
  return (total_tiles + best_per - 1) / best_per;
}

// The arithmetic of plan_tail_helpers (pure: CPU-tested through the test library): `helpers` free CTA pairs for
// `tiles` row tiles whose contractions are nkb0 k-blocks long.  Returns the helper pairs to launch (0 = none) and
// the k-block at which the primaries hand over.
int balance_tail_helpers(int tiles, int helpers, int nkb0, int* kp_out) {
  if (helpers < 1 || nkb0 < 64) return 0;
  helpers = std::min(helpers, tiles);
  const int per_helper = (tiles + helpers - 1) / helpers;  // row tiles whose tail one helper pair takes
  // Balance the primary's kp k-blocks against per_helper tails of nkb0 - kp k-blocks, each of which also costs the
  // helper a pipeline restart, one stored partial tile and a flag: about 8 k-blocks' worth of time per row tile.
  // Measured at 16384^2, K = 256 (nkb0 = 512, 8 tiles per helper): kp = 448 / 456 / 464 / 480 -> 540 / 497 / 493 /
  // 504 us per iteration (the model gives 463); with short contractions (16384 x 2048: nkb0 = 64) the helpers were
  // the slower side and the A GEMM took 69 us instead of 45 - no helpers unless they take at least 8 % off.
  constexpr int kItemOverheadKb = 8;
  int kp = static_cast<int>((static_cast<long long>(per_helper) * (nkb0 + kItemOverheadKb) + per_helper) / (per_helper + 1));
  if (kp * 100LL > nkb0 * 92LL && !std::getenv("NMFB_TAIL_KP")) return 0;
  if (const char* env = std::getenv("NMFB_TAIL_KP")) kp = std::atoi(env);  // tuning experiments
  kp = std::min(std::max(kp, 1), nkb0 - 1);
  *kp_out = kp;
  return helpers;
}

static cudaError_t gemm_attrs_once();

int plan_tail_helpers(const GemmLaunch& L, int epi, int num_sms, int reserve_sms, int* kp_out) {
  if (const char* env = std::getenv("NMFB_TAIL_HELPERS"))
    if (env[0] == '0') return 0;
  const GemmArgs& a = L.args;
  if (L.cg != 2 || L.grid.y != 1 || L.grid.z != 1 || a.nkb_seg != a.nkb0 || a.sk_helpers != 0) return 0;
  if (epi != EPI_STORE && epi != EPI_HUPDATE) return 0;
  const int tiles = static_cast<int>(L.grid.x) / 2;
  if (tiles * 6 <= num_sms) return 0;  // a third of the SMs busy: split-K territory
  // co-resident clusters of this kernel (one CTA per SM: shared memory); the helpers must run beside the primaries.
  // Asked once per process: the occupancy query takes 1 - 100 ms (measured), far too long for every plan.
  static int cached_clusters[2] = {-1, -1};
  static std::mutex cache_mutex;
  int max_clusters = 0;
  if (gemm_attrs_once() != cudaSuccess) return 0;
  {
    std::lock_guard<std::mutex> lock(cache_mutex);
    int& slot = cached_clusters[epi == EPI_STORE ? 0 : 1];
    if (slot < 0) {
      slot = 0;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * 74, 1, 1);
      cfg.blockDim = dim3(kGemmThreads);
      cfg.dynamicSmemBytes = TileCfg<2>::smem_bytes;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int q = 0;
      cudaError_t e = epi == EPI_STORE ? cudaOccupancyMaxActiveClusters(&q, panel_gemm_kernel<EPI_STORE, 2>, &cfg)
                                       : cudaOccupancyMaxActiveClusters(&q, panel_gemm_kernel<EPI_HUPDATE, 2>, &cfg);
      if (e == cudaSuccess) slot = q;
      else cudaGetLastError();
    }
    max_clusters = slot;
  }
  if (max_clusters <= 0) return 0;
  const int pairs = std::min(num_sms / 2, max_clusters) - (reserve_sms + 1) / 2;
  int helpers = pairs - tiles;
  if (const char* env = std::getenv("NMFB_TAIL_HELPERS")) helpers = std::min(helpers, std::atoi(env));
  return balance_tail_helpers(tiles, helpers, a.nkb0, kp_out);
}

void set_tail_helpers(GemmLaunch* L, int helpers, int kp, float* part, unsigned int* flags) {
  GemmArgs& a = L->args;
  a.sk_tiles = static_cast<int>(L->grid.x) / 2;
  a.sk_helpers = helpers;
  a.sk_kp = kp;
  a.sk_part = part;
  a.sk_flag = flags;
  a.sk_epoch = 0;
  L->grid.x = 2 * (a.sk_tiles + helpers);
}

template <int EPI, int CG>
static cudaError_t set_smem_attr() {
  return cudaFuncSetAttribute(panel_gemm_kernel<EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TileCfg<CG>::smem_bytes);
}

template <int EPI>
static cudaError_t launch_one(const GemmLaunch& L, cudaStream_t stream) {
  if (L.cg >= 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = L.grid;
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = TileCfg<2>::smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = L.cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if constexpr (EPI == EPI_STORE || EPI == EPI_HUPDATE) {  // the two kernels of the Euclidean iteration
      if (L.cg == 4)
        return cudaLaunchKernelEx(&cfg, panel_gemm_kernel<EPI, 4>, L.tmX0, L.tmY0, L.tmX1, L.tmY1, L.tmXb, L.tmYb,
                                  L.tmXc, L.tmYc, L.tmH, L.args);
    }
    if (L.cg == 4) return cudaErrorInvalidValue;
    return cudaLaunchKernelEx(&cfg, panel_gemm_kernel<EPI, 2>, L.tmX0, L.tmY0, L.tmX1, L.tmY1, L.tmXb, L.tmYb,
                              L.tmXc, L.tmYc, L.tmH, L.args);
  }
  panel_gemm_kernel<EPI, 1><<<L.grid, kGemmThreads, TileCfg<1>::smem_bytes, stream>>>(
      L.tmX0, L.tmY0, L.tmX1, L.tmY1, L.tmXb, L.tmYb, L.tmXc, L.tmYc, L.tmH, L.args);
  return cudaGetLastError();
}

static cudaError_t gemm_attrs_once() {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&]() {
    cudaError_t e[14] = {set_smem_attr<EPI_STORE, 4>(), set_smem_attr<EPI_HUPDATE, 4>(),
                         set_smem_attr<EPI_ABQ, 1>(),   set_smem_attr<EPI_ABQ, 2>(),
                         set_smem_attr<EPI_STORE, 1>(), set_smem_attr<EPI_HUPDATE, 1>(),
                         set_smem_attr<EPI_RECON, 1>(), set_smem_attr<EPI_RESID, 1>(),
                         set_smem_attr<EPI_KLQ, 1>(),   set_smem_attr<EPI_STORE, 2>(),
                         set_smem_attr<EPI_HUPDATE, 2>(), set_smem_attr<EPI_RECON, 2>(),
                         set_smem_attr<EPI_RESID, 2>(), set_smem_attr<EPI_KLQ, 2>()};
    for (cudaError_t x : e)
      if (x != cudaSuccess) attr_err = x;
  });
  return attr_err;
}

std::string launch_gemm(const GemmLaunch& L, int epi, cudaStream_t stream) {
  const cudaError_t attr_err = gemm_attrs_once();
  if (attr_err != cudaSuccess)
    return std::string("cudaFuncSetAttribute(panel_gemm): ") + cudaGetErrorString(attr_err);
  if (epi != EPI_STORE && L.grid.z != 1) return "launch_gemm: fused epilogues require splits == 1";
  cudaError_t e = cudaSuccess;
  switch (epi) {
    case EPI_STORE: e = launch_one<EPI_STORE>(L, stream); break;
    case EPI_HUPDATE: e = launch_one<EPI_HUPDATE>(L, stream); break;
    case EPI_RECON: e = launch_one<EPI_RECON>(L, stream); break;
    case EPI_RESID: e = launch_one<EPI_RESID>(L, stream); break;
    case EPI_KLQ: e = launch_one<EPI_KLQ>(L, stream); break;
    case EPI_ABQ: e = launch_one<EPI_ABQ>(L, stream); break;
    default: return "launch_gemm: unknown epilogue";
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return std::string("panel_gemm launch: ") + cudaGetErrorString(e);
  return "";
}

}  // namespace nmfb
