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
int set_drop_salt_gemm(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_attn_tc(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_attn_simt(unsigned long long salt, cudaStream_t stream);
int set_drop_salt_elementwise(unsigned long long salt, cudaStream_t stream);

}  // namespace samk

extern "C" {
int samk_version(void) { return 100; }
const char* samk_last_error(void) { return samk::g_err; }
int samk_sm_count(void) { return samk::sm_count(); }
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
