// Error reporting and misc entry points of the samk C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "../../include/samk.h"

namespace samk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}

int sm_count();
int sm_count_physical();
void set_sm_reserved(int n);
int set_drop_salt_gemm(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_attn_tc(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_attn_simt(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_elementwise(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_dev_gemm(const unsigned long long* src, cudaStream_t stream);
int set_drop_salt_dev_attn_tc(const unsigned long long* src, cudaStream_t stream);
int set_drop_salt_dev_attn_simt(const unsigned long long* src, cudaStream_t stream);
int set_drop_salt_dev_elementwise(const unsigned long long* src, cudaStream_t stream);

// [0] = replay counter, [1] = current salt (splitmix64 of the counter)
__device__ unsigned long long g_salt_state[2] = {0ull, 0ull};
__global__ void advance_salt_kernel() {
  unsigned long long z = (g_salt_state[0] += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  g_salt_state[1] = z ^ (z >> 31);
}

}  // namespace samk

extern "C" {
int samk_version(void) { return 100; }
const char* samk_last_error(void) { return samk::g_err; }
int samk_sm_count(void) { return samk::sm_count_physical(); }
int samk_reserve_sms(int n) {
  if (n < 0) { samk::set_error("samk_reserve_sms: negative count"); return SAMK_ERR_ARG; }
  samk::set_sm_reserved(n);
  return samk::sm_count();
}
int samk_advance_dropout_salt(void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* state = nullptr;
  if (cudaGetSymbolAddress((void**)&state, samk::g_salt_state) != cudaSuccess) {
    samk::set_error("samk_advance_dropout_salt: no symbol address: %s", cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  samk::advance_salt_kernel<<<1, 1, 0, s>>>();
  if (samk::set_drop_salt_dev_gemm(state + 1, s) || samk::set_drop_salt_dev_attn_tc(state + 1, s) ||
      samk::set_drop_salt_dev_attn_simt(state + 1, s) || samk::set_drop_salt_dev_elementwise(state + 1, s)) {
    samk::set_error("samk_advance_dropout_salt: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return samk::check_launch("samk_advance_dropout_salt");
}
// Timing events that may be recorded INSIDE a stream capture (cudaEventRecordExternal: the record becomes an event-record
// node of the graph and the event is readable with cudaEventElapsedTime after a replay) -- per-kernel durations of the
// replayed step, which torch.cuda.Event cannot give.
int samk_timing_event_create(void** ev) {
  if (!ev) { samk::set_error("samk_timing_event_create: null pointer"); return SAMK_ERR_ARG; }
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { samk::set_error("cudaEventCreate: %s", cudaGetErrorString(cudaGetLastError())); return SAMK_ERR_CUDA; }
  *ev = (void*)e;
  return SAMK_OK;
}
int samk_timing_event_record(void* ev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &st);
  const unsigned flags = st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
  if (cudaEventRecordWithFlags((cudaEvent_t)ev, s, flags) != cudaSuccess) {
    samk::set_error("cudaEventRecordWithFlags: %s", cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}
int samk_timing_event_elapsed_ms(void* start, void* end, float* ms) {
  if (!ms) { samk::set_error("samk_timing_event_elapsed_ms: null pointer"); return SAMK_ERR_ARG; }
  if (cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)end) != cudaSuccess) {
    samk::set_error("cudaEventElapsedTime: %s", cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}
int samk_timing_event_destroy(void* ev) { return cudaEventDestroy((cudaEvent_t)ev) == cudaSuccess ? SAMK_OK : SAMK_ERR_CUDA; }
int samk_set_dropout_salt(unsigned long long salt, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (samk::set_drop_salt_gemm(salt, s) || samk::set_drop_salt_attn_tc(salt, s) || samk::set_drop_salt_attn_simt(salt, s) ||
      samk::set_drop_salt_elementwise(salt, s)) {
    samk::set_error("samk_set_dropout_salt: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}
}
