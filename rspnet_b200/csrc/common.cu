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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled is not available from this driver (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return encode;
}

static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, CUtensorMapSwizzle swz, const void* base, int rank,
                       const unsigned long long* dims, const unsigned long long* strides_bytes, const unsigned* box,
                       const unsigned* elem_strides = nullptr) {
  EncodeTiledFn encode = tmap_encoder();
  if (!encode) return RSP_ERR_CUDA;
  RSP_REQUIRE(rank >= 1 && rank <= 5, "tensor map: rank %d", rank);
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    RSP_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map: box[%d] = %u out of range", i, box[i]);
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      RSP_REQUIRE(strides_bytes[i - 1] % 16 == 0, "tensor map: stride %d not a multiple of 16 bytes", i);
    }
  }
  RSP_REQUIRE(reinterpret_cast<uintptr_t>(base) % 16 == 0, "tensor map: base address not 16-byte aligned");
  CUresult r = encode(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return RSP_ERR_CUDA;
  }
  return RSP_OK;
}

__global__ void __launch_bounds__(256) zero_fill_kernel(uint4* __restrict__ p16, size_t n16, unsigned char* __restrict__ tail,
                                                        int ntail) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
    p16[i] = make_uint4(0, 0, 0, 0);
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < ntail) tail[threadIdx.x] = 0;
}

cudaError_t zero_async(void* p, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) return cudaMemsetAsync(p, 0, bytes, stream);   // never on the hot path
  const size_t n16 = bytes >> 4;
  const int ntail = static_cast<int>(bytes & 15);
  size_t blocks = (n16 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks == 0) blocks = 1;
  zero_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<uint4*>(p), n16,
                                                                      static_cast<unsigned char*>(p) + (n16 << 4), ntail);
  return cudaGetLastError();
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box) {
  RSP_REQUIRE(box[0] * 2 <= 128, "tensor map: inner box exceeds the 128-byte swizzle span");
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, base, rank, dims, strides_bytes, box);
}

int make_tmap_bf16_strided(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                           const unsigned long long* strides_bytes, const unsigned* box, const unsigned* estr) {
  RSP_REQUIRE(box[0] * 2 <= 128 && estr[0] == 1, "tensor map: inner box exceeds the 128-byte swizzle span");
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, base, rank, dims, strides_bytes, box,
                     estr);
}

// 8-byte elements (one RGBx pixel of four bf16), no swizzle: box rows land back to back, out-of-range elements are zero.
int make_tmap_u64_rows(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                       const unsigned long long* strides_bytes, const unsigned* box) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, CU_TENSOR_MAP_SWIZZLE_NONE, base, rank, dims, strides_bytes, box);
}

// Same with a traversal stride per dimension (box[i] is the traversed extent: box[i] / estr[i] elements are copied).
int make_tmap_u64_rows_strided(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                               const unsigned long long* strides_bytes, const unsigned* box, const unsigned* estr) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, CU_TENSOR_MAP_SWIZZLE_NONE, base, rank, dims, strides_bytes, box,
                     estr);
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
