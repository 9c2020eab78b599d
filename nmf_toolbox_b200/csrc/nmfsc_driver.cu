#include "engine.cuh"
extern "C" int nmfb_nmfsc(nmfb_handle* h, int K, const nmfb_config* cfg, float* W_out, float* H_out,
                          double* cost_out, int* n_cost) {
  return h ? h->fail(NMFB_ERR_UNSUPPORTED, "nmfsc: not built yet") : NMFB_ERR_INVALID_ARGUMENT;
}
