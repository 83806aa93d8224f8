// optim.cu — SURVEY §8(f) N1: the optimizer step for the hash table (and the small parameter groups).
//
// Replaces, per parameter, the eight full-tensor ATen passes of torch::optim::Adam::step
// (torch/csrc/api/src/optim/adam.cpp; the reference builds it at src/ExpRunner.cpp:54 from the groups of
// Hash3DAnchored::OptimParamGroups, src/Field/Hash3DAnchored.cpp:124-150, and steps at ExpRunner.cpp:136) and
// the per-call fp32->fp16 table copy the next forward makes (Hash3DAnchored.cu:186) with ONE pass:
// 16 B read + 12 B written + 2 B fp16 shadow per element.  The update is the same arithmetic in the same
// order with the same roundings as the ATen elementwise kernels (which nvcc contracts to FMAs):
//     g'  = fma(wd, p, g)                        grad.add(p, weight_decay)          (only when wd != 0)
//     m   = fma(1-b1, g', m*b1)                  exp_avg.mul_(b1).add_(g', 1-b1)
//     v   = fma(1-b2, g'*g', v*b2)               exp_avg_sq.mul_(b2).addcmul_(g', g', 1-b2)
//     den = sqrt(v) * float(1/sqrt(bc2)) + eps   (exp_avg_sq.sqrt() / sqrt(bc2)).add_(eps)   [div by a scalar = mul by the
//                                                reciprocal formed in double, as ATen's div_true kernel does]
//     p   = fma(-lr/bc1, m/den, p)               p.addcdiv_(exp_avg, denom, -step_size)
// so a run that swaps the optimizer keeps bit-identical parameters (tests/test_gpu_parity.py::test_fused_adam*).
// Entries beyond `n_live` are never touched by the path (the level-overlap quirk leaves 15/32 of the pool dead:
// zero gradient and zero moments for ever, for which the update is the identity) and are skipped.
#include "common.cuh"

namespace f2b {

struct AdamScalars {
  float b1, one_minus_b1, b2, one_minus_b2, inv_sqrt_bc2, eps, neg_step, wd;
};

template <bool WD, bool SHADOW>
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 __half* __restrict__ shadow, int64_t n4, AdamScalars s) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 P = reinterpret_cast<float4*>(p)[i];
  const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
  float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
  float* pp = reinterpret_cast<float*>(&P);
  const float* gg = reinterpret_cast<const float*>(&G);
  float* mm = reinterpret_cast<float*>(&M);
  float* vv = reinterpret_cast<float*>(&V);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float gr = gg[k];
    if (WD) gr = ffma(s.wd, pp[k], gr);
    mm[k] = ffma(s.one_minus_b1, gr, fmul(mm[k], s.b1));
    vv[k] = ffma(s.one_minus_b2, fmul(gr, gr), fmul(vv[k], s.b2));
    const float den = fadd(fmul(fsqrt(vv[k]), s.inv_sqrt_bc2), s.eps);
    pp[k] = ffma(s.neg_step, fdiv(mm[k], den), pp[k]);
  }
  reinterpret_cast<float4*>(p)[i] = P;
  reinterpret_cast<float4*>(m)[i] = M;
  reinterpret_cast<float4*>(v)[i] = V;
  if (SHADOW) {
    const __half2 h0 = __floats2half2_rn(P.x, P.y), h1 = __floats2half2_rn(P.z, P.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h0);
    o.y = *reinterpret_cast<const uint32_t*>(&h1);
    reinterpret_cast<uint2*>(shadow)[i] = o;
  }
}

}  // namespace f2b

using namespace f2b;

extern "C" int f2b_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t n_live,
                             double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                             void* shadow_f16, void* stream) {
  F2B_REQUIRE(n >= 0 && n_live >= 0 && n_live <= n && step >= 1, "f2b_adam_step: bad sizes / step");
  if (n_live == 0) return F2B_OK;
  F2B_REQUIRE(param && grad && exp_avg && exp_avg_sq, "f2b_adam_step: null pointer");
  F2B_REQUIRE((n_live % 4) == 0, "f2b_adam_step: n_live must be a multiple of 4");
  // scalar preparation exactly as torch::optim::Adam does it in double, then narrowed to the kernels' float
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars s;
  s.b1 = (float)beta1; s.one_minus_b1 = (float)(1.0 - beta1);
  s.b2 = (float)beta2; s.one_minus_b2 = (float)(1.0 - beta2);
  s.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));           // ATen: tensor / Scalar == tensor * float(1.0 / double(scalar))  [probed]
  s.eps = (float)eps;
  s.neg_step = (float)(-(lr / bc1));
  s.wd = (float)weight_decay;
  const int64_t n4 = n_live / 4;
  const unsigned blocks = (unsigned)div_up(n4, 256);
  cudaStream_t st = as_stream(stream);
  __half* sh = (__half*)shadow_f16;
  if (weight_decay != 0.0) {
    if (sh) adam_step_kernel<true, true><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, sh, n4, s);
    else adam_step_kernel<true, false><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, sh, n4, s);
  } else {
    if (sh) adam_step_kernel<false, true><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, sh, n4, s);
    else adam_step_kernel<false, false><<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, sh, n4, s);
  }
  return check_launch("f2b_adam_step");
}
