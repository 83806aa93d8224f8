// runtime.cu — error plumbing, device info and dtype casts of the f2nerf_b200 C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace f2b {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return F2B_ECUDA;
  }
  return F2B_OK;
}

// fp32 -> fp16 with scale (TCNNWP.cpp:111 params cast, :168 dL_doutput*scale cast; tcnn identity
// encoding, encodings/identity.h:45-85) — 128-bit loads, 64-bit stores.
__global__ void cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n,
                                    float scale) {
  const int64_t i4 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i4);
    __half2 lo = __floats2half2_rn(__fmul_rn(v.x, scale), __fmul_rn(v.y, scale));
    __half2 hi = __floats2half2_rn(__fmul_rn(v.z, scale), __fmul_rn(v.w, scale));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + i4) = o;
  } else {
    for (int64_t i = i4; i < n; i++) dst[i] = __float2half_rn(__fmul_rn(src[i], scale));
  }
}
__global__ void cast_f16_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, int64_t n,
                                    float scale) {
  const int64_t i4 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    const uint2 r = *reinterpret_cast<const uint2*>(src + i4);
    const __half2 lo = *reinterpret_cast<const __half2*>(&r.x), hi = *reinterpret_cast<const __half2*>(&r.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    *reinterpret_cast<float4*>(dst + i4) =
        make_float4(__fmul_rn(a.x, scale), __fmul_rn(a.y, scale), __fmul_rn(b.x, scale), __fmul_rn(b.y, scale));
  } else {
    for (int64_t i = i4; i < n; i++) dst[i] = __fmul_rn(__half2float(src[i]), scale);
  }
}
}  // namespace f2b

using namespace f2b;

extern "C" const char* f2b_last_error(void) { return g_err; }
extern "C" int f2b_abi_version(void) { return 1; }

// per-device cache: cudaGetDeviceProperties costs milliseconds and must not sit on the launch path
extern "C" int f2b_device_info(int* sm_count, int* l2_bytes) {
  static int s_sm[64], s_l2[64];
  static bool s_have[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("f2b_device_info: no CUDA device"); return F2B_ECUDA; }
  if (!s_have[dev]) {
    int sm = 0, l2 = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev) != cudaSuccess) {
      set_error("f2b_device_info: query failed");
      return F2B_ECUDA;
    }
    s_sm[dev] = sm; s_l2[dev] = l2; s_have[dev] = true;
  }
  if (sm_count) *sm_count = s_sm[dev];
  if (l2_bytes) *l2_bytes = s_l2[dev];
  return F2B_OK;
}

extern "C" int f2b_cast_f32_to_f16(const float* src, void* dst, int64_t n, float scale, void* stream) {
  if (n <= 0) return F2B_OK;
  F2B_REQUIRE(src && dst, "f2b_cast_f32_to_f16: null pointer");
  cast_f32_f16_kernel<<<div_up(div_up(n, 4), 256), 256, 0, as_stream(stream)>>>(src, (__half*)dst, n, scale);
  return check_launch("f2b_cast_f32_to_f16");
}
extern "C" int f2b_cast_f16_to_f32(const void* src, float* dst, int64_t n, float scale, void* stream) {
  if (n <= 0) return F2B_OK;
  F2B_REQUIRE(src && dst, "f2b_cast_f16_to_f32: null pointer");
  cast_f16_f32_kernel<<<div_up(div_up(n, 4), 256), 256, 0, as_stream(stream)>>>((const __half*)src, dst, n, scale);
  return check_launch("f2b_cast_f16_to_f32");
}
extern "C" int f2b_table_to_half(const float* table_f32, void* table_f16, int64_t n, void* stream) {
  return f2b_cast_f32_to_f16(table_f32, table_f16, n, 1.0f, stream);
}
