// C ABI: handle lifetime, V upload, ReconstructFromDecomposition, projfunc.
// (nmf / cnmf / nmfsc live in their *_driver.cu files.)
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>

#include "comm.cuh"
#include "engine.cuh"
#include "ew_kernels.cuh"

using namespace nmfb;

static std::string g_create_error;

void nmf_session_release(nmfb_handle* h);  // nmf_driver.cu

extern "C" const char* nmfb_version(void) { return "nmfb200 0.1 (sm_100a)"; }

extern "C" const char* nmfb_last_error(const nmfb_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" long long nmfb_launch_count(const nmfb_handle* h) { return h ? h->launches : 0; }
extern "C" int nmfb_last_loop(nmfb_handle* h, int* iters, double* device_ms) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  if (iters) *iters = h->loop_iters;
  if (device_ms) *device_ms = h->loop_ms;
  return NMFB_OK;
}
extern "C" int nmfb_last_halvings(nmfb_handle* h, int* out, int capacity) {
  if (!h) return 0;
  const int n = static_cast<int>(h->halvings.size());
  for (int i = 0; i < n && i < capacity && out; ++i) out[i] = h->halvings[i];
  return n;
}
extern "C" long long nmfb_malloc_count(const nmfb_handle* h) { return h ? h->mallocs : 0; }

extern "C" int nmfb_create(nmfb_handle** out, int device) {
  if (!out) return NMFB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                     "); libnmfb200 has no CPU fallback";
    return NMFB_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    g_create_error = "device index out of range";
    return NMFB_ERR_INVALID_ARGUMENT;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return NMFB_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; libnmfb200 contains sm_100a code only";
    return NMFB_ERR_CUDA;
  }
  e = cudaSetDevice(device);
  nmfb_handle* h = new nmfb_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  {
    const char* np_env = std::getenv("NMFB_NO_POOL");
    h->pool.cap = (np_env && np_env[0] == '1') ? 0 : prop.totalGlobalMem / 3;
  }
  // (stream priorities were tried for the side stream: no gain, slightly slower large GEMMs)
  // (again with tail helpers on 144 SMs, main stream first: no difference)
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e == cudaSuccess) e = cudaMallocHost(&h->pinned, 64 * sizeof(int));
  h->stage_half = size_t(8) << 20;
  if (e == cudaSuccess) e = cudaMallocHost(&h->stage, 2 * h->stage_half);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_stage[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_stage[1], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    g_create_error = std::string("handle creation failed: ") + cudaGetErrorString(e);
    delete h;
    return NMFB_ERR_CUDA;
  }
  *out = h;
  return NMFB_OK;
}

extern "C" void nmfb_destroy(nmfb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  nmf_session_release(h);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  if (h->comm) comm_destroy(h);
  if (h->Vown) dev_free(h, h->Vown, h->Vown_bytes);
  if (h->Vwork) dev_free(h, h->Vwork, h->Vwork_bytes);
  h->pool.trim();
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->stage) cudaFreeHost(h->stage);
  for (cudaEvent_t ev : h->ev_stage)
    if (ev) cudaEventDestroy(ev);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int nmfb_trim(nmfb_handle* h) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  h->pool.trim();
  return NMFB_OK;
}

static void drop_V(nmfb_handle* h) {
  if (h->Vown) {
    cudaStreamSynchronize(h->stream);
    dev_free(h, h->Vown, h->Vown_bytes);
  }
  h->Vown = nullptr;
  h->Vown_bytes = 0;
  h->Vraw = nullptr;
}

extern "C" int nmfb_set_V(nmfb_handle* h, const float* V_host, int m, int n) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  if (!V_host || m <= 0 || n <= 0) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "set_V: bad arguments");
  cudaSetDevice(h->device);
  const bool trace = std::getenv("NMFB_TRACE") != nullptr;
  timespec ts0, ts1, ts2;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  nmf_session_release(h);
  drop_V(h);
  const long long ld = round_up(m, 4);
  const size_t bytes = static_cast<size_t>(n) * ld * sizeof(float);
  void* vp = nullptr;
  NMFB_CUDA(h, dev_alloc(h, &vp, bytes));
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  h->Vown = static_cast<float*>(vp);
  h->Vown_bytes = bytes;
  if (ld != m) NMFB_CUDA(h, cudaMemsetAsync(h->Vown, 0, bytes, h->stream));
  h->Vraw = h->Vown;
  h->m = m;
  h->n = n;
  h->ldv = ld;
  NMFB_TRY(upload_colmajor(h, V_host, m, n, h->Vown, ld));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (trace) {
    clock_gettime(CLOCK_MONOTONIC, &ts2);
    auto ms = [](const timespec& a, const timespec& b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6; };
    fprintf(stderr, "[nmfb] set_V: release+alloc %.1f ms, upload of %.0f MiB %.1f ms (%.1f GB/s), %lld cudaMalloc so far\n",
            ms(ts0, ts1), bytes / 1048576.0, ms(ts1, ts2), bytes / 1e6 / ms(ts1, ts2), h->mallocs);
  }
  return NMFB_OK;
}

extern "C" int nmfb_set_V_device(nmfb_handle* h, const float* V_dev, int m, int n, long long ld) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  if (!V_dev || m <= 0 || n <= 0 || ld < m || ld % 4 != 0 ||
      (reinterpret_cast<uintptr_t>(V_dev) & 15) != 0)
    return h->fail(NMFB_ERR_INVALID_ARGUMENT,
                   "set_V_device: need ld >= m, ld %% 4 == 0 and a 16-byte aligned pointer");
  cudaSetDevice(h->device);
  nmf_session_release(h);
  drop_V(h);
  h->Vraw = V_dev;
  h->m = m;
  h->n = n;
  h->ldv = ld;
  return NMFB_OK;
}

// ------------------------------------------------------------------ ReconstructFromDecomposition
namespace apidetail {
// unrounded shifted stack (RFD.m:37)
__global__ void hstack_raw_kernel(const float* __restrict__ H, float* __restrict__ Hs, int K, int T, int n,
                                  long long ld) {
  const int c = blockIdx.y;
  const int k = c % K, t = c / K;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    Hs[c * ld + j] = (j >= t) ? H[k * ld + (j - t)] : 0.f;
}
}  // namespace apidetail
using namespace apidetail;

// V_hat = W*H (RFD.m:31) or sum_t W(:,:,t)*[zeros(K,t-1) H(:,1:n-t+1)] (RFD.m:33-38).
// The tensor cores multiply tf32 numbers; to return V_hat at fp32 accuracy each
// factor is split into a tf32 head and tail and the contraction is run over the
// stacked operands [W_hi W_lo W_hi] * [H_hi; H_hi; H_lo].
extern "C" int nmfb_reconstruct(nmfb_handle* h, const float* W, const float* H, int m, int K, int T,
                                int n, float* Vhat_out) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  if (!W || !H || !Vhat_out || m <= 0 || K <= 0 || T <= 0 || n <= 0)
    return h->fail(NMFB_ERR_INVALID_ARGUMENT, "reconstruct: bad arguments");
  cudaSetDevice(h->device);
  Arena ar;
  const int KT = K * T, KTp = round_up(KT, 32);
  const long long ldw = round_up(m, 4), ldh = round_up(n, 4);
  float *Wc, *Hm, *Hs, *Xs, *Ys, *out;
  NMFB_TRY(ar.alloc(h, &Wc, static_cast<size_t>(KTp) * ldw));
  NMFB_TRY(ar.alloc(h, &Hm, static_cast<size_t>(K) * ldh));
  NMFB_TRY(ar.alloc(h, &Hs, static_cast<size_t>(KTp) * ldh));
  NMFB_TRY(ar.alloc(h, &Xs, static_cast<size_t>(3) * KTp * ldw));
  NMFB_TRY(ar.alloc(h, &Ys, static_cast<size_t>(3) * KTp * ldh));
  NMFB_TRY(ar.alloc(h, &out, static_cast<size_t>(n) * ldw));
  NMFB_TRY(upload_colmajor(h, W, m, KT, Wc, ldw));
  NMFB_TRY(upload_H(h, &ar, H, K, n, Hm, ldh));
  hstack_raw_kernel<<<vec_grid(n, KT), 256, 0, h->stream>>>(Hm, Hs, K, T, n, ldh);
  NMFB_TRY(check_launch(h, "hstack_raw"));
  const long long cw = static_cast<long long>(KTp) * ldw, ch = static_cast<long long>(KTp) * ldh;
  split_copy_kernel<<<dim3(1024, 1), 256, 0, h->stream>>>(Wc, Xs, Xs + cw, 1, static_cast<int>(cw), cw);
  NMFB_TRY(check_launch(h, "split_tf32(W)"));
  NMFB_CUDA(h, cudaMemcpyAsync(Xs + 2 * cw, Xs, cw * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  split_copy_kernel<<<dim3(1024, 1), 256, 0, h->stream>>>(Hs, Ys, Ys + 2 * ch, 1, static_cast<int>(ch), ch);
  NMFB_TRY(check_launch(h, "split_tf32(H)"));
  NMFB_CUDA(h, cudaMemcpyAsync(Ys + ch, Ys, ch * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  GemmOp op;
  MatRef X{Xs, m, 3 * KTp, ldw, true};
  MatRef Y{Ys, n, 3 * KTp, ldh, true};
  NMFB_TRY(plan_fused(h, &op, EPI_RECON, X, Y, 3 * KTp, nullptr, nullptr, 0, m, round_up(n, 64), n,
                      nullptr));
  op.L.args.Qout = out;
  op.L.args.ldv = ldw;
  NMFB_TRY(run_gemm(h, op));
  NMFB_TRY(download_colmajor(h, out, ldw, m, n, Vhat_out));
  return NMFB_OK;
}

// ------------------------------------------------------------------ projfunc
extern "C" int nmfb_projfunc(nmfb_handle* h, const float* s, int N, int count, double k1, double k2,
                             int nn, float* v_out, int* iters_out) {
  if (!h) return NMFB_ERR_INVALID_ARGUMENT;
  if (!s || !v_out || N <= 0 || count <= 0)
    return h->fail(NMFB_ERR_INVALID_ARGUMENT, "projfunc: bad arguments");
  if (N > kProjThreads * 32 * kProjMaskWords)
    return h->fail(NMFB_ERR_UNSUPPORTED, "projfunc: vectors longer than %d are not supported",
                   kProjThreads * 32 * kProjMaskWords);
  cudaSetDevice(h->device);
  Arena ar;
  float* X;
  int *it, *fail;
  const size_t total = static_cast<size_t>(N) * count;
  NMFB_TRY(ar.alloc(h, &X, total));
  NMFB_TRY(ar.alloc(h, &it, count));
  NMFB_TRY(ar.alloc(h, &fail, 1));
  NMFB_CUDA(h, cudaMemcpyAsync(X, s, total * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  launch_projfunc(h->stream, count, X, N, N, k1, k2, nn, it, fail);
  NMFB_TRY(check_launch(h, "projfunc"));
  int failed = 0;
  NMFB_CUDA(h, cudaMemcpyAsync(v_out, X, total * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaMemcpyAsync(&failed, fail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (iters_out)
    NMFB_CUDA(h, cudaMemcpyAsync(iters_out, it, count * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (failed) return h->fail(NMFB_ERR_PROJFUNC, "projfunc diverged (non-finite values)");
  return NMFB_OK;
}
