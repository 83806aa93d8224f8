// hash.cu — anchored multi-level hash-grid encode (forward gather, backward scatter) for sm_100a.
//
// Replaces Hash3DAnchoredForwardKernel / Hash3DAnchoredBackwardKernel
// (src/Field/Hash3DAnchored.cu:11-155) and the per-call dtype passes of
// Hash3DAnchoredFunction::forward/backward (:160-233) and Hash3DAnchored::AnchoredQuery's
// (pts+1)/2 (src/Field/Hash3DAnchored.cpp:91).
//
// B200 design: one thread owns one sample for all 16 levels (the reference launches 16 grid.y
// passes that each re-read the point), lanes of a warp hold 32 consecutive samples of (mostly)
// one ray so the coarse levels' corners coalesce into the same 32 B sectors in L1; the 8 corner
// loads of a level are issued back to back (8-16 independent 4 B half2 gathers in flight per
// thread), the 32 encoded halfs leave as four 128-bit stores.  The fp16 table (17 MB live at
// log2 19) is L2 resident on B200, so the bound is L2->SM sector throughput, not HBM.
// The blend reproduces the reference build's exact fp32 sequence (read off its PTX):
//   acc = w001*f001; acc = fma(w000,f000,acc); then fma in order 010,011,100,101,110,111.
#include "common.cuh"
#include "hash.cuh"

namespace f2b {

__global__ void __launch_bounds__(128)
hash_fwd_kernel(const __half* __restrict__ table, const int* __restrict__ prim_pool,
                const float* __restrict__ bias_pool, int n_volumes, int local_size,
                const float* __restrict__ pts, const int* __restrict__ vol, int vol_stride, int n_pts,
                __half* __restrict__ out) {
  __shared__ float s_scale[F2B_N_LEVELS];
  if (threadIdx.x < F2B_N_LEVELS) s_scale[threadIdx.x] = level_scale(threadIdx.x);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  uint32_t o[16];
  encode_point(table, prim_pool, bias_pool, n_volumes, local_size, s_scale, __ldg(pts + size_t(i) * 3),
               __ldg(pts + size_t(i) * 3 + 1), __ldg(pts + size_t(i) * 3 + 2),
               __ldg(vol + size_t(i) * vol_stride), o);
  uint4* dst = reinterpret_cast<uint4*>(out + size_t(i) * 32);
#pragma unroll
  for (int q = 0; q < 4; q++) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

// Backward: thread per (sample, level); 8 vector reductions red.global.add.v2.f32 into the fp32
// gradient table (the reference accumulates fp16 atomics of grad*128 and divides later; fp32 is
// both faster on B200's L2 atomic units and more accurate).  Zero-gradient rows are skipped like
// the reference (Hash3DAnchored.cu:149).
template <bool GRAD_F16>
__global__ void __launch_bounds__(256)
hash_bwd_kernel(const int* __restrict__ prim_pool, const float* __restrict__ bias_pool, int n_volumes,
                int local_size, const float* __restrict__ pts, const int* __restrict__ vol,
                int vol_stride, int n_pts, const void* __restrict__ grad_feat, float grad_mul,
                float* __restrict__ grad_table) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int l = int(gid & 15);
  const int64_t i = gid >> 4;
  if (i >= n_pts) return;
  float g0, g1;
  if (GRAD_F16) {
    const __half2 g = reinterpret_cast<const __half2*>(grad_feat)[i * 16 + l];
    const float2 gf = __half22float2(g);
    g0 = gf.x; g1 = gf.y;
  } else {
    const float2 gf = reinterpret_cast<const float2*>(grad_feat)[i * 16 + l];
    g0 = gf.x; g1 = gf.y;
  }
  if (g0 == 0.f && g1 == 0.f) return;
  g0 *= grad_mul; g1 *= grad_mul;
  const float p0 = __ldg(pts + i * 3), p1 = __ldg(pts + i * 3 + 1), p2 = __ldg(pts + i * 3 + 2);
  const float x0 = fmul(fadd(p0, 1.f), .5f), x1 = fmul(fadd(p1, 1.f), .5f), x2 = fmul(fadd(p2, 1.f), .5f);
  const int v = __ldg(vol + i * vol_stride);
  const int tv = l * n_volumes + v;
  Corner8 c;
  corners(x0, x1, x2, level_scale(l), prim_pool + tv * 3, bias_pool + tv * 3, (unsigned)local_size, c);
  float* base = grad_table + size_t(l) * local_size;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    float2* dst = reinterpret_cast<float2*>(base + size_t(c.idx[k]) * 2);
    atomicAdd(dst, make_float2(c.w[k] * g0, c.w[k] * g1));
  }
}

__global__ void level_scales_kernel(float* out) {
  if (threadIdx.x < F2B_N_LEVELS) out[threadIdx.x] = level_scale(threadIdx.x);
}

}  // namespace f2b

using namespace f2b;

extern "C" int f2b_hash_level_scales(float* scales16_host) {
  F2B_REQUIRE(scales16_host, "f2b_hash_level_scales: null pointer");
  float* d = nullptr;
  if (cudaMalloc(&d, 16 * sizeof(float)) != cudaSuccess) { set_error("f2b_hash_level_scales: cudaMalloc failed"); cudaGetLastError(); return F2B_ECUDA; }
  level_scales_kernel<<<1, 32>>>(d);
  cudaError_t e = cudaMemcpy(scales16_host, d, 16 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { set_error("f2b_hash_level_scales: %s", cudaGetErrorString(e)); return F2B_ECUDA; }
  return F2B_OK;
}

extern "C" int f2b_hash_fwd(const void* table_f16, const int* prim_pool, const float* bias_pool,
                            int n_volumes, int local_size, const float* pts, const int* vol,
                            int vol_stride, int n_pts, void* out_f16, void* stream) {
  if (n_pts <= 0) return F2B_OK;
  F2B_REQUIRE(table_f16 && prim_pool && bias_pool && pts && vol && out_f16, "f2b_hash_fwd: null pointer");
  F2B_REQUIRE(n_volumes > 0 && local_size > 0 && (local_size % 2) == 0, "f2b_hash_fwd: bad n_volumes/local_size");
  hash_fwd_kernel<<<div_up(n_pts, 128), 128, 0, as_stream(stream)>>>(
      (const __half*)table_f16, prim_pool, bias_pool, n_volumes, local_size, pts, vol, vol_stride, n_pts,
      (__half*)out_f16);
  return check_launch("f2b_hash_fwd");
}

extern "C" int f2b_hash_bwd(const int* prim_pool, const float* bias_pool, int n_volumes, int local_size,
                            const float* pts, const int* vol, int vol_stride, int n_pts,
                            const void* grad_feat, int grad_is_f16, float grad_mul, float* grad_table,
                            void* stream) {
  if (n_pts <= 0) return F2B_OK;
  F2B_REQUIRE(prim_pool && bias_pool && pts && vol && grad_feat && grad_table, "f2b_hash_bwd: null pointer");
  const int blocks = div_up(int64_t(n_pts) * 16, 256);
  if (grad_is_f16)
    hash_bwd_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(prim_pool, bias_pool, n_volumes, local_size, pts,
                                                                 vol, vol_stride, n_pts, grad_feat, grad_mul, grad_table);
  else
    hash_bwd_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(prim_pool, bias_pool, n_volumes, local_size, pts,
                                                                  vol, vol_stride, n_pts, grad_feat, grad_mul, grad_table);
  return check_launch("f2b_hash_bwd");
}
