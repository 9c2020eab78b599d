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

}  // namespace nmfb
