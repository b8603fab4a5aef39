// Small device helpers shared by the Bartlett kernels.
#pragma once
#include <cuda_runtime.h>

namespace bogp {

// Phase k*r reaches ~1.6e4 rad (388 Hz, 10 km); float32 phase fails the 1e-5 parity bar
// (SURVEY 7.3 item 1), so the product and the reduction to [-pi, pi] are done in float64
// and only the reduced angle goes through the float32 sincos.
__device__ __forceinline__ float reduce_phase(double theta) {
  const double inv2pi = 0.15915494309189533576888;
  const double twopi_hi = 6.283185307179586232;      // fl(2*pi)
  const double twopi_lo = 2.4492935982947064e-16;    // 2*pi - twopi_hi
  double n = rint(theta * inv2pi);
  double r = fma(-n, twopi_hi, theta);
  r = fma(-n, twopi_lo, r);
  return (float)r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Linear interpolation index/weight of the source-depth table (data contract, modes.py).
__device__ __forceinline__ void depth_index(double z, double z0, double dz, int nz_tab, int& i0, float& w) {
  double t = (z - z0) / dz;
  t = fmin(fmax(t, 0.0), (double)(nz_tab - 1));
  int i = (int)floor(t);
  if (i > nz_tab - 2) i = nz_tab - 2;
  i0 = i;
  w = (float)(t - (double)i);
}

__device__ __forceinline__ float combine_tonal(float acc, float bf, int mf, bool first) {
  if (mf == 0) return first ? bf : acc + bf;
  return first ? bf : acc * bf;
}

}  // namespace bogp
