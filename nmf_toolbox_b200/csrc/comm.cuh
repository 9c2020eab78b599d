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
void comm_destroy(Comm* c);

}  // namespace nmfb
