// Error plumbing and device checks for the rspnet_b200 C ABI.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "rspnet_b200.h"

namespace rsp {

static thread_local char g_error[512] = "";
static int g_sm_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return RSP_ERR_CUDA;
  }
  return RSP_OK;
}

int device_sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

}  // namespace rsp

extern "C" {

int rsp_abi_version(void) { return RSP_ABI_VERSION; }

const char* rsp_last_error(void) { return rsp::g_error; }

int rsp_init(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    rsp::set_error("rsp_init: no usable CUDA device: %s", cudaGetErrorString(e));
    return rsp::RSP_ERR_CUDA;
  }
  if (major != 10) {
    rsp::set_error("rsp_init: device has compute capability %d.x; this library is sm_100a only", major);
    return rsp::RSP_ERR_ARCH;
  }
  rsp::g_sm_count = 0;
  rsp::device_sm_count();
  return rsp::RSP_OK;
}

}  // extern "C"
