// Translation unit of tests/libnmfb200_test.so: the panel GEMM with its host-side planner plus the kernel-level
// test hooks.  Test infrastructure only - the product library (libnmfb200.cu) does not contain debug_entry.cu.
#include "gemm_host.cu"
#include "debug_entry.cu"
