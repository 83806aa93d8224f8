// scan.cuh — single-block exclusive/inclusive scan of per-ray counts (shared by sampler.cu and composite.cu).
#pragma once
#include "common.cuh"
namespace f2b {
// counts[R] -> bounds[R,2] = {exclusive, inclusive} prefix (reference: torch::cumsum at :395 and the
// start/end rewrite at :213-216); totals[0] = sum.  One block; R is a few thousand.
static __global__ void __launch_bounds__(1024) scan_counts_kernel(const int* __restrict__ counts, int n,
                                                           int* __restrict__ bounds,
                                                           int* __restrict__ total) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int local = 0;
  for (int i = beg; i < end; i++) local += counts[i];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = s_warp[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += v;
    }
    s_warp[tid] = w;
  }
  __syncthreads();
  int run = incl - local + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
  for (int i = beg; i < end; i++) {
    const int c = counts[i];
    bounds[2 * i] = run;
    run += c;
    bounds[2 * i + 1] = run;
  }
  if (tid == 1023) total[0] = s_warp[31];
}

}  // namespace f2b
