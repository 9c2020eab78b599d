// Single translation unit of libnmfb200.so (all kernels are sm_100a only).
#include "gemm_host.cu"
#include "engine.cu"
#include "comm.cu"
#include "api.cu"
#include "nmf_driver.cu"
#include "debug_entry.cu"
#include "cnmf_driver.cu"
#include "nmfsc_driver.cu"
