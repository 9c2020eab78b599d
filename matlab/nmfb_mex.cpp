// MEX gateway for libnmfb200.so - SOURCE ONLY: this image has no MATLAB (no mex.h, no mex
// compiler), so this file is not built or tested here.  It shows the binding a toolbox
// maintainer adds so that nmf.m / cnmf.m / nmfsc.m / ReconstructFromDecomposition.m /
// projfunc.m keep their signatures and forward to the GPU engine (see matlab/nmf.m etc.).
//
//   [W, H, cost] = nmfb_mex('nmf',   V, K, cfg)        % cfg: struct, fields as nmf.m:17-65
//   [W, H, cost] = nmfb_mex('cnmf',  V, K, T, cfg)
//   [W, H, cost] = nmfb_mex('nmfsc', V, K, cfg)
//   [W, H, Z, cost] = nmfb_mex('constrainednmf', V_ordered, K, col2z, nz, cfg)   % see matlab/constrainednmf.m
//   V_hat        = nmfb_mex('reconstruct', W, H)
//   [v, iters]   = nmfb_mex('projfunc', s, k1, k2, nn)
//
// Build (on a machine with MATLAB + CUDA):  mex -I../include nmfb_mex.cpp -L../nmf_toolbox_b200 -lnmfb200
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"
#include "nmfb200.h"

static nmfb_handle* g_handle = nullptr;

static void cleanup() {
  if (g_handle) nmfb_destroy(g_handle);
  g_handle = nullptr;
}

static nmfb_handle* handle() {
  if (!g_handle) {
    if (nmfb_create(&g_handle, 0) != NMFB_OK) mexErrMsgIdAndTxt("nmfb:create", "%s", nmfb_last_error(nullptr));
    mexAtExit(cleanup);
  }
  return g_handle;
}

static void check(int rc) {
  if (rc == NMFB_OK) return;
  static const char* ids[] = {"nmfb:ok", "nmfb:invalidArgument", "nmfb:cuda", "nmfb:unsupported", "nmfb:divergence",
                              "nmfb:abZero", "nmfb:negativeData", "nmfb:noData", "nmfb:projfunc"};
  mexErrMsgIdAndTxt(ids[rc >= 0 && rc <= 8 ? rc : 1], "%s", nmfb_last_error(g_handle));
}

// MATLAB arrays are column-major; the engine wants column-major float32.
static std::vector<float> to_single(const mxArray* a) {
  const size_t n = mxGetNumberOfElements(a);
  std::vector<float> out(n);
  if (mxIsSingle(a)) {
    std::memcpy(out.data(), mxGetData(a), n * sizeof(float));
  } else if (mxIsDouble(a)) {
    const double* p = mxGetPr(a);
    for (size_t i = 0; i < n; ++i) out[i] = static_cast<float>(p[i]);
  } else {
    mexErrMsgIdAndTxt("nmfb:type", "numeric single/double input expected");
  }
  return out;
}

// Scalar config field.  A one-element cell is a single-source setting (nmf.m:312-359); cells with several
// sources are resolved by the .m wrappers (matlab/nmf.m, matlab/cnmf.m) before the gateway is called and are
// rejected here rather than silently reduced to their first element.
static double field(const mxArray* cfg, const char* name, double dflt) {
  if (!cfg || !mxIsStruct(cfg)) return dflt;
  const mxArray* f = mxGetField(cfg, 0, name);
  if (!f || mxIsEmpty(f)) return dflt;
  if (mxIsCell(f)) {
    if (mxGetNumberOfElements(f) != 1)
      mexErrMsgIdAndTxt("nmfb:invalidArgument", "config.%s: multi-source cell arrays must go through nmf.m / cnmf.m", name);
    f = mxGetCell(f, 0);
  }
  if (!f || mxIsEmpty(f) || mxGetNumberOfElements(f) != 1)
    mexErrMsgIdAndTxt("nmfb:invalidArgument", "config.%s must be a scalar", name);
  return mxGetScalar(f);
}

// Initial factor: single-source cell unwrapped, element count checked against what the engine will read.
static const mxArray* init_field(const mxArray* cfg, const char* name, size_t expect) {
  if (!cfg || !mxIsStruct(cfg)) return nullptr;
  const mxArray* f = mxGetField(cfg, 0, name);
  if (!f || mxIsEmpty(f)) return nullptr;
  if (mxIsCell(f)) {
    if (mxGetNumberOfElements(f) != 1)
      mexErrMsgIdAndTxt("nmfb:invalidArgument", "config.%s: multi-source cell arrays must go through nmf.m / cnmf.m", name);
    f = mxGetCell(f, 0);
    if (!f || mxIsEmpty(f)) return nullptr;
  }
  if (mxGetNumberOfElements(f) != expect)
    mexErrMsgIdAndTxt("nmfb:invalidArgument", "config.%s has %llu elements, expected %llu", name,
                      static_cast<unsigned long long>(mxGetNumberOfElements(f)), static_cast<unsigned long long>(expect));
  return f;
}

static int divergence_code(const mxArray* cfg) {
  if (!cfg || !mxIsStruct(cfg)) return NMFB_DIV_EUCLIDEAN;
  const mxArray* f = mxGetField(cfg, 0, "divergence");
  if (!f) return NMFB_DIV_EUCLIDEAN;  // nmf.m:250-252
  char buf[64];
  mxGetString(f, buf, sizeof(buf));
  const std::string s(buf);
  if (s == "euclidean") return NMFB_DIV_EUCLIDEAN;
  if (s == "kl_divergence" || s == "kl") return NMFB_DIV_KL;
  if (s == "frobenius") return NMFB_DIV_FROBENIUS;
  if (s == "is_divergence" || s == "is") return NMFB_DIV_IS;
  if (s == "ab_divergence" || s == "ab") return NMFB_DIV_AB;
  return 99;  // -> NMFB_ERR_DIVERGENCE, nmf.m:165-166
}

static mxArray* from_single(const std::vector<float>& v, mwSize r, mwSize c, mwSize t = 1) {
  mwSize dims[3] = {r, c, t};
  mxArray* a = mxCreateNumericArray(t > 1 ? 3 : 2, dims, mxSINGLE_CLASS, mxREAL);
  std::memcpy(mxGetData(a), v.data(), v.size() * sizeof(float));
  return a;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("nmfb:usage", "first argument: command string");
  char cmdbuf[32];
  mxGetString(prhs[0], cmdbuf, sizeof(cmdbuf));
  const std::string cmd(cmdbuf);
  nmfb_handle* h = handle();

  if (cmd == "nmf" || cmd == "lnmf" || cmd == "cnmf" || cmd == "nmfsc" || cmd == "cnmfsc") {
    const bool conv = cmd == "cnmf" || cmd == "cnmfsc";
    const mxArray* V = prhs[1];
    const int m = static_cast<int>(mxGetM(V)), n = static_cast<int>(mxGetN(V));
    if (nrhs < (conv ? 4 : 3)) mexErrMsgIdAndTxt("nmfb:usage", "%s: too few arguments", cmd.c_str());
    if (mxIsCell(prhs[2]) && mxGetNumberOfElements(prhs[2]) != 1)
      mexErrMsgIdAndTxt("nmfb:invalidArgument", "num_basis_elems: multi-source cell arrays must go through nmf.m / cnmf.m");
    const mxArray* Karg = mxIsCell(prhs[2]) ? mxGetCell(prhs[2], 0) : prhs[2];
    if (!Karg || mxGetNumberOfElements(Karg) != 1) mexErrMsgIdAndTxt("nmfb:invalidArgument", "num_basis_elems must be a scalar");
    const int K = static_cast<int>(mxGetScalar(Karg));
    const int T = conv ? static_cast<int>(mxGetScalar(prhs[3])) : 1;
    if (K <= 0 || T <= 0 || m <= 0 || n <= 0) mexErrMsgIdAndTxt("nmfb:invalidArgument", "sizes must be positive");
    const mxArray* cfg = nrhs > (conv ? 4 : 3) ? prhs[conv ? 4 : 3] : nullptr;
    std::vector<float> Vf = to_single(V), W0, H0;
    check(nmfb_set_V(h, Vf.data(), m, n));
    nmfb_config c;
    std::memset(&c, 0, sizeof(c));
    c.divergence = (cmd == "nmfsc" || cmd == "cnmfsc" || cmd == "lnmf") ? 0 : divergence_code(cfg);
    c.alpha = field(cfg, "alpha", 1);
    c.beta = field(cfg, "beta", 1);
    c.W_sparsity = field(cfg, "W_sparsity", 0);
    c.H_sparsity = field(cfg, "H_sparsity", 0);
    c.W_fixed = field(cfg, "W_fixed", 0) != 0;
    c.H_fixed = field(cfg, "H_fixed", 0) != 0;
    c.maxiter = static_cast<int>(field(cfg, "maxiter", 0));
    c.tolerance = field(cfg, "tolerance", 0);
    {  // wrong-shaped initial factors would make the engine read past the host arrays
      const mxArray* w = init_field(cfg, "W_init", static_cast<size_t>(m) * K * T);
      const mxArray* hh = init_field(cfg, "H_init", static_cast<size_t>(K) * n);
      if (w) { W0 = to_single(w); c.W_init = W0.data(); }
      if (hh) { H0 = to_single(hh); c.H_init = H0.data(); }
    }
    // per-basis vectors (what nmf.m's per-source cell settings become after concatenating the
    // sources, see matlab/nmf.m): numeric vectors with one entry per basis column
    std::vector<double> lwk, lhk;
    std::vector<int> fwk, fhk;
    auto vec_field = [&](const char* name, std::vector<double>* d, std::vector<int>* i) {
      const mxArray* f = (cfg && mxIsStruct(cfg)) ? mxGetField(cfg, 0, name) : nullptr;
      if (!f || mxIsEmpty(f)) return false;
      if (static_cast<int>(mxGetNumberOfElements(f)) != K)
        mexErrMsgIdAndTxt("nmfb:config", "%s needs one value per basis", name);
      const mxArray* dbl = f;
      mxArray* conv_arr = nullptr;
      if (!mxIsDouble(f)) {
        mxArray* in = const_cast<mxArray*>(f);
        mexCallMATLAB(1, &conv_arr, 1, &in, "double");
        dbl = conv_arr;
      }
      const double* p = mxGetPr(dbl);
      for (int k = 0; k < K; ++k) {
        if (d) d->push_back(p[k]);
        if (i) i->push_back(p[k] != 0);
      }
      if (conv_arr) mxDestroyArray(conv_arr);
      return true;
    };
    if (cmd == "nmf") {
      if (vec_field("W_sparsity_k", &lwk, nullptr)) c.W_sparsity_k = lwk.data();
      if (vec_field("H_sparsity_k", &lhk, nullptr)) c.H_sparsity_k = lhk.data();
      if (vec_field("W_fixed_k", nullptr, &fwk)) c.W_fixed_k = fwk.data();
      if (vec_field("H_fixed_k", nullptr, &fhk)) c.H_fixed_k = fhk.data();
    }
    const int maxiter = c.maxiter > 0 ? c.maxiter : 100;
    std::vector<float> W(static_cast<size_t>(m) * K * T), H(static_cast<size_t>(K) * n);
    std::vector<double> cost(maxiter + 1);
    int ncost = 0;
    if (cmd == "nmf") check(nmfb_nmf(h, K, &c, W.data(), H.data(), cost.data(), &ncost));
    else if (cmd == "lnmf") check(nmfb_lnmf(h, K, &c, W.data(), H.data(), cost.data(), &ncost));
    else if (cmd == "cnmfsc") check(nmfb_cnmfsc(h, K, T, &c, W.data(), H.data(), cost.data(), &ncost));
    else if (conv) check(nmfb_cnmf(h, K, T, &c, W.data(), H.data(), cost.data(), &ncost));
    else check(nmfb_nmfsc(h, K, &c, W.data(), H.data(), cost.data(), &ncost));
    plhs[0] = from_single(W, m, K, T);
    if (nlhs > 1) plhs[1] = from_single(H, K, n);
    if (nlhs > 2) {
      plhs[2] = mxCreateDoubleMatrix(ncost, 1, mxREAL);  // trimmed like nmf.m:222
      std::memcpy(mxGetPr(plhs[2]), cost.data(), ncost * sizeof(double));
    }
  } else if (cmd == "constrainednmf") {
    // [W, H, Z, cost] = nmfb_mex('constrainednmf', V_ordered, K, col2z (int32, n x 1), nz, cfg)
    if (nrhs < 5) mexErrMsgIdAndTxt("nmfb:usage", "constrainednmf: V, K, col2z, nz[, cfg]");
    const mxArray* V = prhs[1];
    const int m = static_cast<int>(mxGetM(V)), n = static_cast<int>(mxGetN(V));
    const int K = static_cast<int>(mxGetScalar(prhs[2]));
    const int nz = static_cast<int>(mxGetScalar(prhs[4]));
    if (!mxIsInt32(prhs[3]) || static_cast<int>(mxGetNumberOfElements(prhs[3])) != n)
      mexErrMsgIdAndTxt("nmfb:invalidArgument", "col2z must be an int32 vector with one entry per sample");
    if (K <= 0 || nz <= 0 || m <= 0 || n <= 0) mexErrMsgIdAndTxt("nmfb:invalidArgument", "sizes must be positive");
    const mxArray* cfg = nrhs > 5 ? prhs[5] : nullptr;
    std::vector<float> Vf = to_single(V), W0, Z0;
    check(nmfb_set_V(h, Vf.data(), m, n));
    nmfb_config c;
    std::memset(&c, 0, sizeof(c));
    c.divergence = divergence_code(cfg);
    c.alpha = field(cfg, "alpha", 1);
    c.beta = field(cfg, "beta", 1);
    c.W_sparsity = field(cfg, "W_sparsity", 0);
    c.H_sparsity = field(cfg, "H_sparsity", 0);  // Z_sparsity (renamed by constrainednmf.m)
    c.W_fixed = field(cfg, "W_fixed", 0) != 0;
    c.H_fixed = field(cfg, "H_fixed", 0) != 0;   // Z_fixed
    c.maxiter = static_cast<int>(field(cfg, "maxiter", 0));
    c.tolerance = field(cfg, "tolerance", 0);
    const mxArray* w = init_field(cfg, "W_init", static_cast<size_t>(m) * K);
    const mxArray* z = init_field(cfg, "Z_init", static_cast<size_t>(K) * nz);
    if (w) { W0 = to_single(w); c.W_init = W0.data(); }
    if (z) Z0 = to_single(z);
    const int maxiter = c.maxiter > 0 ? c.maxiter : 100;
    std::vector<float> W(static_cast<size_t>(m) * K), H(static_cast<size_t>(K) * n), Z(static_cast<size_t>(K) * nz);
    std::vector<double> cost(maxiter + 1);
    int ncost = 0;
    check(nmfb_constrainednmf(h, K, &c, static_cast<const int*>(mxGetData(prhs[3])), nz, z ? Z0.data() : nullptr, W.data(),
                              H.data(), Z.data(), cost.data(), &ncost));
    plhs[0] = from_single(W, m, K);
    if (nlhs > 1) plhs[1] = from_single(H, K, n);
    if (nlhs > 2) plhs[2] = from_single(Z, K, nz);
    if (nlhs > 3) {
      plhs[3] = mxCreateDoubleMatrix(ncost, 1, mxREAL);
      std::memcpy(mxGetPr(plhs[3]), cost.data(), ncost * sizeof(double));
    }
  } else if (cmd == "reconstruct") {
    const mxArray* W = prhs[1];
    const mxArray* H = prhs[2];
    const mwSize* d = mxGetDimensions(W);
    const int m = static_cast<int>(d[0]), K = static_cast<int>(d[1]);
    const int T = mxGetNumberOfDimensions(W) == 3 ? static_cast<int>(d[2]) : 1;
    const int n = static_cast<int>(mxGetN(H));
    std::vector<float> Wf = to_single(W), Hf = to_single(H), out(static_cast<size_t>(m) * n);
    check(nmfb_reconstruct(h, Wf.data(), Hf.data(), m, K, T, n, out.data()));
    plhs[0] = from_single(out, m, n);
  } else if (cmd == "projfunc") {
    std::vector<float> s = to_single(prhs[1]), v(s.size());
    int iters = 0;
    check(nmfb_projfunc(h, s.data(), static_cast<int>(s.size()), 1, mxGetScalar(prhs[2]), mxGetScalar(prhs[3]),
                        mxGetScalar(prhs[4]) != 0, v.data(), &iters));
    plhs[0] = from_single(v, s.size(), 1);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(iters);
  } else {
    mexErrMsgIdAndTxt("nmfb:usage", "unknown command %s", cmd.c_str());
  }
}
