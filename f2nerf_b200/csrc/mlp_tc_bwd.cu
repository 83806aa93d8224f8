// mlp_tc_bwd.cu — MLP backward on tcgen05 tensor cores (impl 1).
//
// Replaces kernel_mlp_fused_backward (External/tiny-cuda-nn/src/fully_fused_mlp.cu:150-259), the CUTLASS
// dL/dinput GEMM (:832-835) and the three split-K weight-gradient GEMMs on side streams (:785,819,828)
// with ONE persistent kernel per call.  Per 128-sample tile (thread i == sample i == TMEM lane i):
//   activation gradients  dH = dOut.Wout, dH0 = dH1.Wh, dIn = dH0.W0   — UMMA M128, A = K-major tile of the
//       gradient rows, B = the weight matrix used as MN-major operand straight from its row-major layout
//       (no transposed weight copies);
//   weight gradients      dW = dAct^T . Act  — batch-reduction UMMAs with K = the 128 samples of the tile:
//       both operands MN-major views of the very same shared-memory tiles, M = 64 output rows, fp32
//       accumulators that stay resident in TMEM across all tiles of the CTA and are flushed once with
//       atomicAdd at the end (no split-K workspace, no fp16 accumulation, no reduction kernels).
// TMEM columns: [0,64) activation-gradient accumulator, then dW0 (32), dWout^T (16), dWh (64 when NH).
#include "common.cuh"
#include "tc.cuh"

namespace f2b {
using namespace tc;

constexpr int kBT = 128;   // tile rows

template <int NH>
struct BwdSmem {
  static constexpr int W0 = 0;                       // [64 x 32]  SW64   4 KB   (B operand, MN-major, K = out row)
  static constexpr int WH = 4096;                    // [64 x 64]  SW128  8 KB
  static constexpr int WO = 12288;                   // [16 x 64]  SW128  2 KB
  static constexpr int DO = 14336;                   // [128 x 16] SW32   4 KB
  static constexpr int X = 18432;                    // [128 x 32] SW64   8 KB
  static constexpr int H0 = 26624;                   // [128 x 64] SW128 16 KB
  static constexpr int DH0 = 43008;                  // [128 x 64] SW128 16 KB
  static constexpr int H1 = 59392;                   // NH only
  static constexpr int DH1 = 75776;                  // NH only
  static constexpr int BAR = NH ? 92160 : 59392;
  static constexpr int BYTES = BAR + 64 + 1024;
  static constexpr int TMEM_COLS = NH ? 256 : 128;
  static constexpr int C_ACT = 0, C_GW0 = 64, C_GWO = 96, C_GWH = 128;
};

__device__ __forceinline__ uint32_t pk(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// MN-major operand: tile of [k rows][row_bytes] with matching swizzle; k-step = 16 rows
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile, int row_bytes, int k_step) {
  const uint64_t layout = row_bytes == 128 ? kLayoutSW128 : (row_bytes == 64 ? kLayoutSW64 : kLayoutSW32);
  return make_desc(tile + k_step * 16 * row_bytes, 16, 8 * row_bytes, layout);
}

template <int K>
__device__ __forceinline__ void stage_w(const __half* __restrict__ w, int rows, unsigned char* dst) {
  constexpr int chunks = K / 8;
  for (int i = threadIdx.x; i < rows * chunks; i += blockDim.x) {
    const int r = i / chunks, c = i % chunks;
    *reinterpret_cast<uint4*>(dst + (K == 64 ? sw128_off(r, c) : sw64_off(r, c))) = *reinterpret_cast<const uint4*>(w + r * K + c * 8);
  }
}

// masked epilogue: acc (64 fp32 from TMEM) * [h > 0] -> fp16 row of the dH tile (SW128)
__device__ __forceinline__ void relu_bwd_epilogue(uint32_t tmem_row, const unsigned char* h_tile, unsigned char* dh_tile, int row) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t r[16];
    tmem_ld16(tmem_row + 16 * q, r);
    tmem_ld_wait();
    uint4 hv[2], o[2];
    hv[0] = *reinterpret_cast<const uint4*>(h_tile + sw128_off(row, 2 * q));
    hv[1] = *reinterpret_cast<const uint4*>(h_tile + sw128_off(row, 2 * q + 1));
    const __half2* hh = reinterpret_cast<const __half2*>(hv);
    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const float2 hf = __half22float2(hh[e]);
      ow[e] = pk(hf.x > 0.f ? __uint_as_float(r[2 * e]) : 0.f, hf.y > 0.f ? __uint_as_float(r[2 * e + 1]) : 0.f);
    }
    *reinterpret_cast<uint4*>(dh_tile + sw128_off(row, 2 * q)) = o[0];
    *reinterpret_cast<uint4*>(dh_tile + sw128_off(row, 2 * q + 1)) = o[1];
  }
}

template <int NH>
__global__ void __launch_bounds__(kBT)
mlp_bwd_tc_kernel(const __half* __restrict__ dout, const __half* __restrict__ in, const __half* __restrict__ hidden,
                  const __half* __restrict__ params, int n_pts, __half* __restrict__ din, float* __restrict__ dparams) {
  using S = BwdSmem<NH>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + S::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + S::BAR + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  stage_w<32>(params, 64, sm + S::W0);
  if (NH) stage_w<64>(params + 64 * 32, 64, sm + S::WH);
  stage_w<64>(params + 64 * 32 + NH * 64 * 64, 16, sm + S::WO);
  if (tid == 0) mbar_init(mbar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, S::TMEM_COLS);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + (uint32_t(warp * 32) << 16);
  const uint32_t s_w0 = smem_u32(sm + S::W0), s_wh = smem_u32(sm + S::WH), s_wo = smem_u32(sm + S::WO);
  const uint32_t s_do = smem_u32(sm + S::DO), s_x = smem_u32(sm + S::X), s_h0 = smem_u32(sm + S::H0);
  const uint32_t s_dh0 = smem_u32(sm + S::DH0), s_h1 = smem_u32(sm + S::H1), s_dh1 = smem_u32(sm + S::DH1);
  unsigned char* const h_last = sm + (NH ? S::H1 : S::H0);
  unsigned char* const dh_last = sm + (NH ? S::DH1 : S::DH0);
  const uint32_t s_hl = NH ? s_h1 : s_h0, s_dhl = NH ? s_dh1 : s_dh0;
  // instruction descriptors
  constexpr uint32_t id_act64 = idesc_f16_f32(128, 64, 0, 1);     // A K-major, B MN-major
  constexpr uint32_t id_act32 = idesc_f16_f32(128, 32, 0, 1);
  constexpr uint32_t id_gw16 = idesc_f16_f32(64, 16, 1, 1);       // both MN-major
  constexpr uint32_t id_gw32 = idesc_f16_f32(64, 32, 1, 1);
  constexpr uint32_t id_gw64 = idesc_f16_f32(64, 64, 1, 1);
  uint32_t phase = 0, first = 1;

  const int n_tiles = (n_pts + kBT - 1) / kBT;
  const __half* hid_last = hidden + size_t(NH) * n_pts * 64;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p = tile * kBT + tid;
    const bool valid = p < n_pts;
    const uint4 z = make_uint4(0, 0, 0, 0);
    {   // stage this sample's rows: dOut (32 B), x (64 B), h0 (128 B), h1 (128 B)
      const uint4* s = reinterpret_cast<const uint4*>(dout + size_t(p) * 16);
#pragma unroll
      for (int c = 0; c < 2; c++) *reinterpret_cast<uint4*>(sm + S::DO + sw32_off(tid, c)) = valid ? __ldg(s + c) : z;
      s = reinterpret_cast<const uint4*>(in + size_t(p) * 32);
#pragma unroll
      for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(sm + S::X + sw64_off(tid, c)) = valid ? __ldg(s + c) : z;
      s = reinterpret_cast<const uint4*>(hidden + size_t(p) * 64);
#pragma unroll
      for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(sm + S::H0 + sw128_off(tid, c)) = valid ? __ldg(s + c) : z;
      if (NH) {
        s = reinterpret_cast<const uint4*>(hid_last + size_t(p) * 64);
#pragma unroll
        for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(sm + S::H1 + sw128_off(tid, c)) = valid ? __ldg(s + c) : z;
      }
    }
    fence_async_smem();
    __syncthreads();
    // ---- A: dH_last = dOut . Wout  (K = 16) -----------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
      mma_f16(tmem + S::C_ACT, kmajor_desc(s_do, 32), mnmajor_desc(s_wo, 128, 0), id_act64, 0);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_sync();
    relu_bwd_epilogue(tmem_row + S::C_ACT, h_last, dh_last, tid);
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (NH) {
      // ---- B: dH0 = dH1 . Wh ; dWout^T += H1^T . dOut ; dWh += dH1^T . H0 -------------------------
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; k++)
          mma_f16(tmem + S::C_ACT, kmajor_desc(s_dh1 + 32 * k, 128), mnmajor_desc(s_wh, 128, k), id_act64, k);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          mma_f16(tmem + S::C_GWO, mnmajor_desc(s_h1, 128, k), mnmajor_desc(s_do, 32, k), id_gw16, (k > 0) | (first ^ 1));
          mma_f16(tmem + S::C_GWH, mnmajor_desc(s_dh1, 128, k), mnmajor_desc(s_h0, 128, k), id_gw64, (k > 0) | (first ^ 1));
        }
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase); phase ^= 1;
      fence_after_sync();
      relu_bwd_epilogue(tmem_row + S::C_ACT, sm + S::H0, sm + S::DH0, tid);
      fence_before_sync();
      fence_async_smem();
      __syncthreads();
    }
    // ---- C: dIn = dH0 . W0 ; dW0 += dH0^T . X ; (NH == 0: dWout^T += H0^T . dOut) ---------------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; k++)
        mma_f16(tmem + S::C_ACT, kmajor_desc(s_dh0 + 32 * k, 128), mnmajor_desc(s_w0, 64, k), id_act32, k);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        mma_f16(tmem + S::C_GW0, mnmajor_desc(s_dh0, 128, k), mnmajor_desc(s_x, 64, k), id_gw32, (k > 0) | (first ^ 1));
        if (!NH) mma_f16(tmem + S::C_GWO, mnmajor_desc(s_hl, 128, k), mnmajor_desc(s_do, 32, k), id_gw16, (k > 0) | (first ^ 1));
      }
      mma_commit(mbar);
    }
    first = 0;
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_sync();
    {
#pragma unroll
      for (int q = 0; q < 2; q++) {
        uint32_t r[16];
        tmem_ld16(tmem_row + S::C_ACT + 16 * q, r);
        tmem_ld_wait();
        if (din && valid) {
          uint4 o[2];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
          for (int e = 0; e < 8; e++) ow[e] = pk(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
          uint4* dst = reinterpret_cast<uint4*>(din + size_t(p) * 32) + 2 * q;
          dst[0] = o[0]; dst[1] = o[1];
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    (void)s_dhl;
  }
  // ---- flush the weight-gradient accumulators (M = 64 layout: warp w, lanes 0..15 hold rows 16w..16w+15) ----
  if (!first) {
    fence_after_sync();
    const int row = warp * 16 + lane;
    float* g0 = dparams;
    float* gh = dparams + 64 * 32;
    float* go = gh + NH * 64 * 64;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      uint32_t r[16];
      tmem_ld16(tmem_row + S::C_GW0 + 16 * q, r);
      tmem_ld_wait();
      if (lane < 16)
#pragma unroll
        for (int e = 0; e < 16; e++) atomicAdd(g0 + row * 32 + 16 * q + e, __uint_as_float(r[e]));
    }
    {
      uint32_t r[16];
      tmem_ld16(tmem_row + S::C_GWO, r);
      tmem_ld_wait();
      if (lane < 16)
#pragma unroll
        for (int e = 0; e < 16; e++) atomicAdd(go + e * 64 + row, __uint_as_float(r[e]));      // transposed back
    }
    if (NH) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t r[16];
        tmem_ld16(tmem_row + S::C_GWH + 16 * q, r);
        tmem_ld_wait();
        if (lane < 16)
#pragma unroll
          for (int e = 0; e < 16; e++) atomicAdd(gh + row * 64 + 16 * q + e, __uint_as_float(r[e]));
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, S::TMEM_COLS);
}

}  // namespace f2b

using namespace f2b;

extern "C" int f2b_mlp_bwd_tc(const void* dout_f16, const void* in_f16, const void* hidden_save_f16,
                              const void* params_f16, int n_hidden_matmuls, int n_pts, void* din_f16,
                              float* dparams_f32, void* stream) {
  if (n_pts <= 0) return F2B_OK;
  F2B_REQUIRE(dout_f16 && in_f16 && hidden_save_f16 && params_f16 && dparams_f32, "f2b_mlp_bwd: null pointer");
  F2B_REQUIRE(n_hidden_matmuls == 0 || n_hidden_matmuls == 1, "f2b_mlp_bwd: n_hidden_matmuls must be 0 or 1");
  int sms = 148;
  f2b_device_info(&sms, nullptr);
  const int n_tiles = div_up(n_pts, kBT);
  if (n_hidden_matmuls == 0) {
    const int grid = n_tiles < sms * 3 ? n_tiles : sms * 3;
    cudaFuncSetAttribute(mlp_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<0>::BYTES);
    mlp_bwd_tc_kernel<0><<<grid, kBT, BwdSmem<0>::BYTES, as_stream(stream)>>>(
        (const __half*)dout_f16, (const __half*)in_f16, (const __half*)hidden_save_f16, (const __half*)params_f16, n_pts,
        (__half*)din_f16, dparams_f32);
  } else {
    const int grid = n_tiles < sms * 2 ? n_tiles : sms * 2;
    cudaFuncSetAttribute(mlp_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<1>::BYTES);
    mlp_bwd_tc_kernel<1><<<grid, kBT, BwdSmem<1>::BYTES, as_stream(stream)>>>(
        (const __half*)dout_f16, (const __half*)in_f16, (const __half*)hidden_save_f16, (const __half*)params_f16, n_pts,
        (__half*)din_f16, dparams_f32);
  }
  return check_launch("f2b_mlp_bwd(tcgen05)");
}
