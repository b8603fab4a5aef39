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

// sin / cos of a LARGE float64 phase for kernels that are bound by instruction issue: reduction modulo pi/2 in
// float64 (quadrant = n mod 4), then the cephes single-precision kernels on |r| <= pi/4 (~1e-7 absolute) -- about
// half the instructions of reduce_phase + sincosf, which reduces a second time in float32.
__device__ __forceinline__ void sincos_big(double theta, float& s, float& c) {
  const double two_over_pi = 0.63661977236758134308;
  const double pio2_hi = 1.57079632679489655800;       // fl(pi/2)
  const double pio2_lo = 6.12323399573676603587e-17;   // pi/2 - pio2_hi
  const double n = rint(theta * two_over_pi);
  double r = fma(-n, pio2_hi, theta);
  r = fma(-n, pio2_lo, r);
  const int q = (int)(long long)n & 3;
  const float x = (float)r, z = x * x;
  const float sr = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, x, x);
  const float cr = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f) * z, z,
                        fmaf(-0.5f, z, 1.0f));
  const float a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;      // sin, cos before the sign
  s = (q & 2) ? -a : a;
  c = ((q + 1) & 2) ? -b : b;
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
