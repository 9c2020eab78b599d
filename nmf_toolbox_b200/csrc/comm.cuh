// Multi-GPU plumbing: one process per GPU, one NCCL communicator per handle.
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy torch already
// mapped into the process when the Python layer is used, else the system one),
// so a single-GPU user needs no NCCL at all.  The only collective on the data
// path is the per-iteration all-reduce of the packed W-step inputs
// (SURVEY.md section 8e): issued as one NCCL group on the handle's stream.
#pragma once
#include <cstddef>

#include "engine.cuh"

namespace nmfb {

int comm_size(const Comm* c);
int comm_rank(const Comm* c);
// In-place sum over ranks of up to three buffers (any may be null / empty), one group call.
int comm_allreduce(nmfb_handle* h, float* f, size_t nf, double* d1, size_t n1, double* d2, size_t n2);
void comm_destroy(nmfb_handle* h);  // also releases the shared region

// Peer-memory path (NVLink, CUDA IPC).  The communicator owns ONE device allocation per rank, mapped
// into every peer process, that holds whatever a session all-reduces (same size and layout on every
// rank; a session asks for it with comm_acquire_region and gets zeroed memory; the allocation and its
// mappings are kept across sessions because cudaIpcOpenMemHandle costs ~100 ms).  comm_allreduce() on
// pointers inside the region runs as our own kernels over peer memory instead of NCCL:
//   barrier -> every rank reduces its slice of the fp32 part reading all peers and writes the sums
//   into all peers (two-shot all-reduce) and sums the few fp64 scalars -> barrier -> scalars in place.
// If IPC is unavailable on any rank (or NMFB_NO_P2P is set) the region is plain memory and NCCL runs.
int comm_acquire_region(nmfb_handle* h, size_t bytes, char** data);

// ---- building blocks for kernels that talk to the peers themselves (row-sharded W step, w_shard.cuh)
constexpr int kMaxRanks = 8;
constexpr int kMaxBlocks = 256;                                      // grid cap of kernels with per-block barriers
constexpr size_t kFlagBytes = kMaxBlocks * kMaxRanks * sizeof(int);  // one barrier site: flags[block][rank]
// Barrier sites (each with its own monotone epoch sequence): 0/1 all-reduce open/close, 2..4 sharded W step
// (4 closing, 6 opening), 5 row gather of the fp32 W at the end of a run
constexpr int kBarrierSites = 9;  // (site 7: cnmf halo exchange, site 8: teardown handshake)
constexpr size_t kHeader = kBarrierSites * kFlagBytes;  // region = [flags | data]
struct PeerTable {
  char* base[kMaxRanks];  // base[q] = rank q's registered region as seen from this process
  int rank, nranks;
};
// true + table when the peer-memory path is active on this communicator
bool comm_peer_table(const nmfb_handle* h, PeerTable* out);
// first of `count` fresh values of the (per communicator) epoch counter; every launch of a kernel that
// uses barrier sites takes fresh epochs so that flags only ever grow (same sequence on every rank)
int comm_next_epoch(nmfb_handle* h, int count = 1);
// offset of a pointer inside the local region (for addressing the same bytes in a peer's mapping)
size_t comm_region_offset(const nmfb_handle* h, const void* p);

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Block b of every rank meets block b of all other ranks (flags[b][rank] in each rank's region
// header, monotone epochs).  The release store of the signalling thread follows the bar.sync, so it
// publishes the whole block's earlier writes (and, by stream order, those of earlier kernels).
__device__ __forceinline__ void p2p_block_barrier(const PeerTable& t, size_t flags_off, int epoch) {
  __syncthreads();
  if (threadIdx.x < t.nranks) {
    const int q = threadIdx.x;
    const size_t row = flags_off + static_cast<size_t>(blockIdx.x) * kMaxRanks * sizeof(int);
    st_release_sys(reinterpret_cast<int*>(t.base[q] + row) + t.rank, epoch);
    const int* mine = reinterpret_cast<const int*>(t.base[t.rank] + row) + q;
    long long t0 = clock64();
    while (ld_acquire_sys(mine) - epoch < 0) {
      if (clock64() - t0 > 30000000000LL) {
        printf("nmfb: peer barrier timeout (rank %d block %d waiting for rank %d, epoch %d, site %d)\n", t.rank,
               blockIdx.x, q, epoch, static_cast<int>(flags_off / kFlagBytes));
        __trap();
      }
    }
  }
  __syncthreads();
}
#endif

}  // namespace nmfb
