// Tensor-core (tcgen05 / TMEM / TMA) path of the engine — DSG_PRECISION_BF16.
#include "dsg_engine.h"

int dsg_tc_create(dsg_engine* e) {
  (void)e;
  return dsg_fail(DSG_ERR_UNSUPPORTED, "DSG_PRECISION_BF16 is not built yet");
}
void dsg_tc_destroy(dsg_engine* e) { (void)e; }
int dsg_tc_denoise(dsg_engine*, int, const float*, const int*, StepRef, float*, cudaStream_t) {
  return dsg_fail(DSG_ERR_UNSUPPORTED, "DSG_PRECISION_BF16 is not built yet");
}
int dsg_tc_run_steps(dsg_engine*, int, float*, int, int, uint64_t, int, cudaStream_t) {
  return dsg_fail(DSG_ERR_UNSUPPORTED, "DSG_PRECISION_BF16 is not built yet");
}
