// Library-level entry points of the C ABI (include/ezflow_b200.h).
#include "ezb_common.cuh"

extern "C" int ezb_version(void) { return EZB_VERSION; }

extern "C" int ezb_last_error(char* buf, size_t n) {
  const char* e = ezb::err_buf();
  const size_t len = strlen(e);
  if (buf && n) {
    const size_t k = len < n - 1 ? len : n - 1;
    memcpy(buf, e, k);
    buf[k] = 0;
  }
  return (int)len;
}

extern "C" int ezb_event_create(void** ev) {
  EZB_REQUIRE(ev, "null pointer");
  cudaEvent_t e;
  EZB_CUDA_OK(cudaEventCreate(&e));
  *ev = e;
  return EZB_OK;
}
extern "C" int ezb_event_destroy(void* ev) {
  if (ev) EZB_CUDA_OK(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
  return EZB_OK;
}
extern "C" int ezb_event_elapsed_ms(void* ev_start, void* ev_end, float* ms) {
  EZB_REQUIRE(ev_start && ev_end && ms, "null pointer");
  EZB_CUDA_OK(cudaEventSynchronize(static_cast<cudaEvent_t>(ev_end)));
  EZB_CUDA_OK(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(ev_start), static_cast<cudaEvent_t>(ev_end)));
  return EZB_OK;
}
