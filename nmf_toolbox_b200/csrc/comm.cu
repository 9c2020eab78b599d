#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <vector>

namespace nmfb {

namespace ncclbind {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) {
      api.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(api.lib, name);
      if (!p && api.err.empty()) api.err = std::string("libnccl.so.2 lacks ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}
}  // namespace ncclbind
using namespace ncclbind;

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-memory region (comm_register_region)
  bool p2p = false;
  PeerTable table{};
  char* region = nullptr;   // allocation owned by the communicator: [flags 256 B | data]
  size_t region_bytes = 0;  // capacity incl. the flag header
  int epoch = 0;
};

// ---------------------------------------------------------------- peer-memory all-reduce kernels
template <int N>
__device__ __forceinline__ void p2p_reduce_slice(const PeerTable& t, size_t f_off, size_t lo, size_t hi) {
  // kU x N independent float4 per thread (64 registers) keep ~10 MB of reads in flight per GPU
  constexpr int kU = 16 / N;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i0 = lo + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i0 < hi; i0 += stride * kU) {
    float4 v[kU][N];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const size_t i = i0 + u * stride;
#pragma unroll
      for (int q = 0; q < N; ++q)
        v[u][q] = i < hi ? __ldcv(reinterpret_cast<const float4*>(t.base[q] + f_off) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const size_t i = i0 + u * stride;
      float4 acc = v[u][0];
#pragma unroll
      for (int q = 1; q < N; ++q) {  // fixed rank order; every element is summed by exactly one rank
        acc.x += v[u][q].x;
        acc.y += v[u][q].y;
        acc.z += v[u][q].z;
        acc.w += v[u][q].w;
      }
      if (i < hi) {
#pragma unroll
        for (int q = 0; q < N; ++q) reinterpret_cast<float4*>(t.base[q] + f_off)[i] = acc;
      }
    }
  }
}

// One kernel per all-reduce.  Blocks never wait for other blocks of their own GPU: when the kernel
// has finished on a rank, every block has passed its closing barrier, hence every block of every
// rank has finished reading and writing - which is all the next kernel in the stream needs.
__global__ void __launch_bounds__(512)
p2p_allreduce_kernel(PeerTable t, int epoch, size_t f_off, size_t nf, size_t d1_off, int n1, size_t d2_off, int n2) {
  p2p_block_barrier(t, 0, epoch);  // every rank's partial results are complete and visible
  const int N = t.nranks;
  // fp32 part: rank r owns float4 slice r
  const size_t n4 = (nf + 3) / 4;
  const size_t per = (n4 + N - 1) / N;
  const size_t lo = per * t.rank, hi = min(n4, lo + per);
  switch (N) {
    case 2: p2p_reduce_slice<2>(t, f_off, lo, hi); break;
    case 3: p2p_reduce_slice<3>(t, f_off, lo, hi); break;
    case 4: p2p_reduce_slice<4>(t, f_off, lo, hi); break;
    case 5: p2p_reduce_slice<5>(t, f_off, lo, hi); break;
    case 6: p2p_reduce_slice<6>(t, f_off, lo, hi); break;
    case 7: p2p_reduce_slice<7>(t, f_off, lo, hi); break;
    default: p2p_reduce_slice<8>(t, f_off, lo, hi); break;
  }
  // fp64 scalars (block 0 only): every rank forms the same sums in rank order, keeps them in
  // registers across the closing barrier and then overwrites its own copy
  double dsum[2] = {0.0, 0.0};
  if (blockIdx.x == 0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = threadIdx.x + u * blockDim.x;
      if (i < n1 + n2) {
        const size_t off = i < n1 ? d1_off + static_cast<size_t>(i) * 8 : d2_off + static_cast<size_t>(i - n1) * 8;
        for (int q = 0; q < N; ++q) dsum[u] += __ldcv(reinterpret_cast<const double*>(t.base[q] + off));
      }
    }
  }
  p2p_block_barrier(t, kFlagBytes, epoch);  // all reads of our data and writes into it are done
  if (blockIdx.x == 0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = threadIdx.x + u * blockDim.x;
      if (i < n1 + n2) {
        const size_t off = i < n1 ? d1_off + static_cast<size_t>(i) * 8 : d2_off + static_cast<size_t>(i - n1) * 8;
        *reinterpret_cast<double*>(t.base[t.rank] + off) = dsum[u];
      }
    }
  }
}

bool comm_peer_table(const nmfb_handle* h, PeerTable* out) {
  const Comm* c = h->comm;
  if (!c || !c->p2p || c->nranks <= 1) return false;
  if (out) *out = c->table;
  return true;
}
int comm_next_epoch(nmfb_handle* h, int count) {
  const int first = h->comm->epoch + 1;
  h->comm->epoch += count;
  return first;
}
size_t comm_region_offset(const nmfb_handle* h, const void* p) {
  return static_cast<size_t>(static_cast<const char*>(p) - h->comm->region);
}

int comm_size(const Comm* c) { return c ? c->nranks : 1; }
int comm_rank(const Comm* c) { return c ? c->rank : 0; }

static void region_release(nmfb_handle* h, Comm* c, bool collective);

void comm_destroy(nmfb_handle* h) {
  Comm* c = h->comm;
  if (!c) return;
  h->comm = nullptr;
  // not a collective: handles may be destroyed at unrelated times on the ranks.  Peers stopped
  // touching our region at the barrier that ended their last all-reduce; they only hold a mapping.
  region_release(h, c, false);
  NcclApi* api = nccl_api();
  if (c->comm && api->CommDestroy) api->CommDestroy(c->comm);
  delete c;
}

static bool in_region(const Comm* c, const void* p, size_t bytes) {
  const char* b = c->region;
  const char* q = static_cast<const char*>(p);
  return q >= b && q + bytes <= b + c->region_bytes;
}

int comm_allreduce(nmfb_handle* h, float* f, size_t nf, double* d1, size_t n1, double* d2, size_t n2) {
  Comm* c = h->comm;
  if (!c || c->nranks <= 1) return NMFB_OK;
  if (!f) nf = 0;
  if (!d1) n1 = 0;
  if (!d2) n2 = 0;
  if (c->p2p && (nf == 0 || (in_region(c, f, nf * 4) && (reinterpret_cast<uintptr_t>(f) & 15) == 0)) &&
      (n1 == 0 || in_region(c, d1, n1 * 8)) && (n2 == 0 || in_region(c, d2, n2 * 8)) && n1 + n2 <= 1024) {
    const char* b = c->region;
    const size_t f_off = nf ? reinterpret_cast<const char*>(f) - b : 0;
    const size_t d1_off = n1 ? reinterpret_cast<const char*>(d1) - b : 0;
    const size_t d2_off = n2 ? reinterpret_cast<const char*>(d2) - b : 0;
    const size_t n4 = (nf + 3) / 4 / c->nranks;
    // one co-resident block per SM at most: blocks spin on their peers
    const int cap = std::min(h->num_sms, kMaxBlocks);
    const int blocks = static_cast<int>(std::max<size_t>(1, std::min<size_t>(cap, (n4 + 2047) / 2048)));
    const int e = ++c->epoch;
    p2p_allreduce_kernel<<<blocks, 512, 0, h->stream>>>(c->table, e, f_off, nf, d1_off, static_cast<int>(n1), d2_off,
                                                         static_cast<int>(n2));
    h->launches += 1;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "peer all-reduce launch failed: %s", cudaGetErrorString(ce));
    return NMFB_OK;
  }
  NcclApi* api = nccl_api();
  ncclResult_t r = api->GroupStart();
  if (r == ncclSuccess && f && nf) r = api->AllReduce(f, f, nf, ncclFloat32, ncclSum, c->comm, h->stream);
  if (r == ncclSuccess && d1 && n1) r = api->AllReduce(d1, d1, n1, ncclFloat64, ncclSum, c->comm, h->stream);
  if (r == ncclSuccess && d2 && n2) r = api->AllReduce(d2, d2, n2, ncclFloat64, ncclSum, c->comm, h->stream);
  ncclResult_t r2 = api->GroupEnd();
  if (r == ncclSuccess) r = r2;
  ++h->launches;
  if (r != ncclSuccess) return h->fail(NMFB_ERR_CUDA, "ncclAllReduce failed: %s", api->GetErrorString(r));
  return NMFB_OK;
}

static void region_release(nmfb_handle* h, Comm* c, bool collective) {
  if (!c->region) return;
  cudaStreamSynchronize(h->stream);
  bool peers_gone = true;
  if (c->p2p) {
    // Freeing memory that a peer still has mapped is undefined.  Handles are destroyed at unrelated times, so the
    // teardown cannot be a collective; instead every rank tells its peers "I am closing my mapping of your region"
    // (a flag in THEIR region header, written just before the close) and frees its own region only once all
    // peers have said so - or, if a peer never does (it may have died), after a bounded wait it leaves the region
    // to the process teardown rather than freeing it under a live mapping.
    const size_t off = 8 * kFlagBytes;
    const int closing = 1;
    for (int q = 0; q < c->nranks; ++q)
      if (q != c->rank && c->table.base[q])
        cudaMemcpy(c->table.base[q] + off + c->rank * sizeof(int), &closing, sizeof(int), cudaMemcpyHostToDevice);
    for (int q = 0; q < c->nranks; ++q)
      if (q != c->rank && c->table.base[q]) cudaIpcCloseMemHandle(c->table.base[q]);
    c->p2p = false;
    if (!collective) {
      int flags[kMaxRanks];
      timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      for (;;) {
        peers_gone = cudaMemcpy(flags, c->region + off, sizeof(flags), cudaMemcpyDeviceToHost) == cudaSuccess;
        for (int q = 0; peers_gone && q < c->nranks; ++q)
          if (q != c->rank && flags[q] == 0) peers_gone = false;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (peers_gone || (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9 > 2.0) break;
        timespec nap{0, 2000000};
        nanosleep(&nap, nullptr);
      }
      cudaGetLastError();
    }
    // nobody may free its region while a peer still has it mapped: meet once through NCCL
    float* one = nullptr;
    if (collective && cudaMalloc(&one, 4) == cudaSuccess) {
      cudaMemsetAsync(one, 0, 4, h->stream);
      nccl_api()->AllReduce(one, one, 1, ncclFloat32, ncclSum, c->comm, h->stream);
      cudaStreamSynchronize(h->stream);
      cudaFree(one);
    }
  }
  if (peers_gone) cudaFree(c->region);  // else: still mapped somewhere - left to the process teardown
  c->region = nullptr;
  c->region_bytes = 0;
}

// Try to map every peer's region (all ranks decide together); on any failure the NCCL path stays.
static int region_export(nmfb_handle* h, Comm* c) {
  if (std::getenv("NMFB_NO_P2P") || c->nranks > kMaxRanks) return NMFB_OK;
  NcclApi* api = nccl_api();
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  bool ok = cudaIpcGetMemHandle(&mine, c->region) == cudaSuccess;
  if (!ok) cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  char *dsend = nullptr, *drecv = nullptr;
  NMFB_CUDA(h, cudaMalloc(&dsend, 64));
  NMFB_CUDA(h, cudaMalloc(&drecv, 64 * c->nranks));
  NMFB_CUDA(h, cudaMemcpyAsync(dsend, &mine, 64, cudaMemcpyHostToDevice, h->stream));
  ncclResult_t r = api->AllGather(dsend, drecv, 64, ncclChar, c->comm, h->stream);
  std::vector<cudaIpcMemHandle_t> all(c->nranks);
  cudaError_t ce = cudaMemcpyAsync(all.data(), drecv, 64 * c->nranks, cudaMemcpyDeviceToHost, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  cudaFree(dsend);
  cudaFree(drecv);
  if (r != ncclSuccess || ce != cudaSuccess) return h->fail(NMFB_ERR_CUDA, "exchange of IPC handles failed");
  // a rank whose export failed sent zeros; everyone sees that and skips the mapping
  const cudaIpcMemHandle_t zero{};
  for (int q = 0; q < c->nranks; ++q)
    if (std::memcmp(&all[q], &zero, sizeof(zero)) == 0) ok = false;
  for (int q = 0; q < c->nranks; ++q) {
    c->table.base[q] = nullptr;
    if (q == c->rank) {
      c->table.base[q] = c->region;
    } else if (ok) {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
      }
      c->table.base[q] = static_cast<char*>(p);
    }
  }
  // all ranks must agree on the path: sum of failures through NCCL
  float* flag = nullptr;
  NMFB_CUDA(h, cudaMalloc(&flag, 4));
  const float mineok = ok ? 0.f : 1.f;
  NMFB_CUDA(h, cudaMemcpyAsync(flag, &mineok, 4, cudaMemcpyHostToDevice, h->stream));
  api->AllReduce(flag, flag, 1, ncclFloat32, ncclSum, c->comm, h->stream);
  float bad = 0.f;
  NMFB_CUDA(h, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, h->stream));
  NMFB_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(flag);
  c->table.rank = c->rank;
  c->table.nranks = c->nranks;
  c->epoch = 0;
  if (bad != 0.f) {
    for (int q = 0; q < c->nranks; ++q)
      if (q != c->rank && c->table.base[q]) cudaIpcCloseMemHandle(c->table.base[q]);
    return NMFB_OK;  // NCCL path
  }
  c->p2p = true;
  return NMFB_OK;
}

int comm_acquire_region(nmfb_handle* h, size_t bytes, char** data) {
  Comm* c = h->comm;
  *data = nullptr;
  if (!c) return h->fail(NMFB_ERR_INVALID_ARGUMENT, "comm_acquire_region without a communicator");
  bytes = (bytes + 255) / 256 * 256;
  if (c->region && c->region_bytes >= bytes + kHeader) {
    // reuse: the mapping, the flags and the epoch counter carry on; only the data is cleared.  Safe
    // against slower peers because our last all-reduce ended with a barrier behind their last access.
    NMFB_CUDA(h, cudaMemsetAsync(c->region + kHeader, 0, bytes, h->stream));
    *data = c->region + kHeader;
    return NMFB_OK;
  }
  region_release(h, c, true);  // every rank grows its region in the same session setup
  NMFB_CUDA(h, cudaMalloc(&c->region, bytes + kHeader));
  c->region_bytes = bytes + kHeader;
  NMFB_CUDA(h, cudaMemsetAsync(c->region, 0, bytes + kHeader, h->stream));
  NMFB_TRY(region_export(h, c));
  *data = c->region + kHeader;
  return NMFB_OK;
}

}  // namespace nmfb

extern "C" int nmfb_comm_unique_id(char id_out[NMFB_UNIQUE_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= NMFB_UNIQUE_ID_BYTES, "unique id does not fit");
  nmfb::NcclApi* api = nmfb::nccl_api();
  if (!api->err.empty() || !id_out) return NMFB_ERR_CUDA;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return NMFB_ERR_CUDA;
  std::memset(id_out, 0, NMFB_UNIQUE_ID_BYTES);
  std::memcpy(id_out, &id, sizeof(id));
  return NMFB_OK;
}

extern "C" int nmfb_comm_init(nmfb_handle* h, const char id[NMFB_UNIQUE_ID_BYTES], int rank, int nranks) {
  if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return NMFB_ERR_INVALID_ARGUMENT;
  nmfb::NcclApi* api = nmfb::nccl_api();
  if (!api->err.empty()) return h->fail(NMFB_ERR_CUDA, "%s", api->err.c_str());
  cudaSetDevice(h->device);
  if (h->comm) {
    nmfb::comm_destroy(h);
  }
  if (nranks == 1) return NMFB_OK;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  nmfb::Comm* c = new nmfb::Comm();
  c->rank = rank;
  c->nranks = nranks;
  ncclResult_t r = api->CommInitRank(&c->comm, nranks, uid, rank);
  if (r != ncclSuccess) {
    delete c;
    return h->fail(NMFB_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r));
  }
  h->comm = c;
  return NMFB_OK;
}
