#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

namespace nmfb {

namespace ncclbind {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
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
};

int comm_size(const Comm* c) { return c ? c->nranks : 1; }
int comm_rank(const Comm* c) { return c ? c->rank : 0; }

void comm_destroy(Comm* c) {
  if (!c) return;
  NcclApi* api = nccl_api();
  if (c->comm && api->CommDestroy) api->CommDestroy(c->comm);
  delete c;
}

int comm_allreduce(nmfb_handle* h, float* f, size_t nf, double* d1, size_t n1, double* d2, size_t n2) {
  Comm* c = h->comm;
  if (!c || c->nranks <= 1) return NMFB_OK;
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
    nmfb::comm_destroy(h->comm);
    h->comm = nullptr;
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
