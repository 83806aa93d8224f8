// mlp_tc.cu — the fully-fused MLP on Blackwell's 5th-generation tensor cores (impl 1).
//
// Replaces tiny-cuda-nn's wmma 16x16x16 fp16-accumulate kernels (kernel_mlp_fused /
// kernel_mlp_fused_backward, External/tiny-cuda-nn/src/fully_fused_mlp.cu:150-259,499-557) with
// tcgen05.mma (UMMA M128 x N{64,16} x K16, fp16 operands, fp32 accumulators in TMEM):
//   * one CTA (4 warps) owns a 128-sample tile; thread i == sample i == TMEM lane i;
//   * the layer input is written to shared memory in the canonical K-major swizzled layout
//     (64 B rows / SWIZZLE_64B for K=32, 128 B rows / SWIZZLE_128B for K=64), the weights sit
//     resident in shared memory in the same layouts for the CTA's lifetime (persistent tiles);
//   * one elected thread issues K/16 MMAs per layer and commits to an mbarrier; all 128 threads then
//     pull their accumulator row out of TMEM with tcgen05.ld (32x32b), apply ReLU, round to fp16 and
//     write the next layer's operand row straight back to shared memory — activations never visit HBM
//     except for the optional hidden_save the backward pass needs;
//   * TMEM: 64 columns per CTA (the 16-column output accumulator reuses the hidden one), so up to
//     8 CTAs share an SM and overlap one tile's global loads with another's MMA/epilogue.
#include "common.cuh"
#include "tc.cuh"

namespace f2b {
using namespace tc;

constexpr int kTcTile = 128;
constexpr int kTmemCols = 64;

struct TcSmem {                         // offsets from a 1024-byte aligned base
  static constexpr int A0 = 0;          // [128 x 32] f16, SW64   (8 KB)
  static constexpr int A1 = 8192;       // [128 x 64] f16, SW128  (16 KB)
  static constexpr int W0 = 24576;      // [64 x 32]  f16, SW64   (4 KB)
  static constexpr int WH = 28672;      // [64 x 64]  f16, SW128  (8 KB)
  static constexpr int WO = 36864;      // [16 x 64]  f16, SW128  (2 KB)
  static constexpr int BAR = 38912;     // mbarrier (8 B) + tmem base (4 B)
  static constexpr int BYTES = 38912 + 64 + 1024;   // + alignment slack
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// weights [rows x K] fp16 row-major (global) -> swizzled K-major tile in shared memory
template <int K>
__device__ __forceinline__ void stage_weights(const __half* __restrict__ w, int rows, unsigned char* dst) {
  constexpr int chunks = K / 8;                                  // 16-byte chunks per row
  for (int i = threadIdx.x; i < rows * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i % chunks;
    const uint4 v = *reinterpret_cast<const uint4*>(w + r * K + c * 8);
    *reinterpret_cast<uint4*>(dst + (K == 64 ? sw128_off(r, c) : sw64_off(r, c))) = v;
  }
}

template <int NH>
__global__ void __launch_bounds__(kTcTile)
mlp_fwd_tc_kernel(const __half* __restrict__ in, const __half* __restrict__ params, int n_pts,
                  __half* __restrict__ out, __half* __restrict__ hidden_save) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + TcSmem::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TcSmem::BAR + 8);
  const int tid = threadIdx.x, warp = tid >> 5;

  stage_weights<32>(params, 64, sm + TcSmem::W0);
  if (NH) stage_weights<64>(params + 64 * 32, 64, sm + TcSmem::WH);
  stage_weights<64>(params + 64 * 32 + NH * 64 * 64, 16, sm + TcSmem::WO);
  if (tid == 0) mbar_init(mbar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + (uint32_t(warp * 32) << 16);   // this warp's 32 TMEM lanes

  const uint32_t a0 = smem_u32(sm + TcSmem::A0), a1 = smem_u32(sm + TcSmem::A1);
  const uint32_t w0 = smem_u32(sm + TcSmem::W0), wh = smem_u32(sm + TcSmem::WH), wo = smem_u32(sm + TcSmem::WO);
  constexpr uint32_t idesc64 = idesc_f16_f32(128, 64), idesc16 = idesc_f16_f32(128, 16);
  uint32_t phase = 0;

  const int n_tiles = (n_pts + kTcTile - 1) / kTcTile;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p = tile * kTcTile + tid;
    const bool valid = p < n_pts;
    // ---- stage the input row (64 B) into A0 -------------------------------------------------
    {
      const uint4* src = reinterpret_cast<const uint4*>(in + size_t(p) * 32);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const uint4 v = valid ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sm + TcSmem::A0 + sw64_off(tid, c)) = v;
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- layer 0: D[128x64] = A0[128x32] . W0^T ------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 2; k++)
        mma_f16(tmem, kmajor_desc(a0 + 32 * k, 64), kmajor_desc(w0 + 32 * k, 64), idesc64, k);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_sync();
    // epilogue: ReLU, fp16, write my row of A1 (and the hidden save)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      uint32_t r[16];
      tmem_ld16(tmem_row + 16 * q, r);
      tmem_ld_wait();
      uint4 v0, v1;
      v0.x = pack_half2(fmaxf(__uint_as_float(r[0]), 0.f), fmaxf(__uint_as_float(r[1]), 0.f));
      v0.y = pack_half2(fmaxf(__uint_as_float(r[2]), 0.f), fmaxf(__uint_as_float(r[3]), 0.f));
      v0.z = pack_half2(fmaxf(__uint_as_float(r[4]), 0.f), fmaxf(__uint_as_float(r[5]), 0.f));
      v0.w = pack_half2(fmaxf(__uint_as_float(r[6]), 0.f), fmaxf(__uint_as_float(r[7]), 0.f));
      v1.x = pack_half2(fmaxf(__uint_as_float(r[8]), 0.f), fmaxf(__uint_as_float(r[9]), 0.f));
      v1.y = pack_half2(fmaxf(__uint_as_float(r[10]), 0.f), fmaxf(__uint_as_float(r[11]), 0.f));
      v1.z = pack_half2(fmaxf(__uint_as_float(r[12]), 0.f), fmaxf(__uint_as_float(r[13]), 0.f));
      v1.w = pack_half2(fmaxf(__uint_as_float(r[14]), 0.f), fmaxf(__uint_as_float(r[15]), 0.f));
      *reinterpret_cast<uint4*>(sm + TcSmem::A1 + sw128_off(tid, 2 * q)) = v0;
      *reinterpret_cast<uint4*>(sm + TcSmem::A1 + sw128_off(tid, 2 * q + 1)) = v1;
      if (hidden_save && valid) {
        uint4* dst = reinterpret_cast<uint4*>(hidden_save + size_t(p) * 64) + 2 * q;
        dst[0] = v0; dst[1] = v1;
      }
    }
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (NH) {
      // ---- hidden layer: D[128x64] = A1[128x64] . Wh^T, result back into A1 -------------------------
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; k++)
          mma_f16(tmem, kmajor_desc(a1 + 32 * k, 128), kmajor_desc(wh + 32 * k, 128), idesc64, k);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase); phase ^= 1;
      fence_after_sync();
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t r[16];
        tmem_ld16(tmem_row + 16 * q, r);
        tmem_ld_wait();
        uint4 v0, v1;
        v0.x = pack_half2(fmaxf(__uint_as_float(r[0]), 0.f), fmaxf(__uint_as_float(r[1]), 0.f));
        v0.y = pack_half2(fmaxf(__uint_as_float(r[2]), 0.f), fmaxf(__uint_as_float(r[3]), 0.f));
        v0.z = pack_half2(fmaxf(__uint_as_float(r[4]), 0.f), fmaxf(__uint_as_float(r[5]), 0.f));
        v0.w = pack_half2(fmaxf(__uint_as_float(r[6]), 0.f), fmaxf(__uint_as_float(r[7]), 0.f));
        v1.x = pack_half2(fmaxf(__uint_as_float(r[8]), 0.f), fmaxf(__uint_as_float(r[9]), 0.f));
        v1.y = pack_half2(fmaxf(__uint_as_float(r[10]), 0.f), fmaxf(__uint_as_float(r[11]), 0.f));
        v1.z = pack_half2(fmaxf(__uint_as_float(r[12]), 0.f), fmaxf(__uint_as_float(r[13]), 0.f));
        v1.w = pack_half2(fmaxf(__uint_as_float(r[14]), 0.f), fmaxf(__uint_as_float(r[15]), 0.f));
        // the MMA that read A1 has completed (mbarrier), so the row can be overwritten in place
        *reinterpret_cast<uint4*>(sm + TcSmem::A1 + sw128_off(tid, 2 * q)) = v0;
        *reinterpret_cast<uint4*>(sm + TcSmem::A1 + sw128_off(tid, 2 * q + 1)) = v1;
        if (hidden_save && valid) {
          uint4* dst = reinterpret_cast<uint4*>(hidden_save + size_t(n_pts) * 64 + size_t(p) * 64) + 2 * q;
          dst[0] = v0; dst[1] = v1;
        }
      }
      fence_before_sync();
      fence_async_smem();
      __syncthreads();
    }
    // ---- output layer: D[128x16] = A1[128x64] . Wout^T (linear) ------------------------------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; k++)
        mma_f16(tmem, kmajor_desc(a1 + 32 * k, 128), kmajor_desc(wo + 32 * k, 128), idesc16, k);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_sync();
    {
      uint32_t r[16];
      tmem_ld16(tmem_row, r);
      tmem_ld_wait();
      if (valid) {
        uint4 v0, v1;
        v0.x = pack_half2(__uint_as_float(r[0]), __uint_as_float(r[1]));   v0.y = pack_half2(__uint_as_float(r[2]), __uint_as_float(r[3]));
        v0.z = pack_half2(__uint_as_float(r[4]), __uint_as_float(r[5]));   v0.w = pack_half2(__uint_as_float(r[6]), __uint_as_float(r[7]));
        v1.x = pack_half2(__uint_as_float(r[8]), __uint_as_float(r[9]));   v1.y = pack_half2(__uint_as_float(r[10]), __uint_as_float(r[11]));
        v1.z = pack_half2(__uint_as_float(r[12]), __uint_as_float(r[13])); v1.w = pack_half2(__uint_as_float(r[14]), __uint_as_float(r[15]));
        uint4* dst = reinterpret_cast<uint4*>(out + size_t(p) * 16);
        dst[0] = v0; dst[1] = v1;
      }
    }
    fence_before_sync();
    __syncthreads();          // TMEM columns and A0/A1 are free for the next tile
  }
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace f2b

using namespace f2b;

extern "C" int f2b_mlp_fwd_tc(const void* in_f16, const void* params_f16, int n_hidden_matmuls, int n_pts,
                              void* out_f16, void* hidden_save_f16, void* stream) {
  if (n_pts <= 0) return F2B_OK;
  F2B_REQUIRE(in_f16 && params_f16 && out_f16, "f2b_mlp_fwd: null pointer");
  F2B_REQUIRE(n_hidden_matmuls == 0 || n_hidden_matmuls == 1, "f2b_mlp_fwd: n_hidden_matmuls must be 0 or 1");
  int sms = 148;
  f2b_device_info(&sms, nullptr);
  const int n_tiles = div_up(n_pts, kTcTile);
  const int grid = n_tiles < sms * 5 ? n_tiles : sms * 5;
  if (n_hidden_matmuls == 0) {
    cudaFuncSetAttribute(mlp_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::BYTES);
    mlp_fwd_tc_kernel<0><<<grid, kTcTile, TcSmem::BYTES, as_stream(stream)>>>((const __half*)in_f16, (const __half*)params_f16,
                                                                              n_pts, (__half*)out_f16, (__half*)hidden_save_f16);
  } else {
    cudaFuncSetAttribute(mlp_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::BYTES);
    mlp_fwd_tc_kernel<1><<<grid, kTcTile, TcSmem::BYTES, as_stream(stream)>>>((const __half*)in_f16, (const __half*)params_f16,
                                                                              n_pts, (__half*)out_f16, (__half*)hidden_save_f16);
  }
  return check_launch("f2b_mlp_fwd(tcgen05)");
}
