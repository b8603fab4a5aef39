// Latency of one hand-shake hop between two warps of a CTA, by ping-pong: (0) mbarrier arrive + try_wait, (1) mbarrier
// arrive + test_wait spin, (2) st.release.cta + ld.acquire.cta spin on a shared-memory word, (3) volatile st / ld spin,
// (4) one signaller -> 16 waiting warps (mbarrier try_wait), all 16 arrive back on a count-16 mbarrier,
// (5) the same through shared-memory words (ld.acquire spin; red.release.add back).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_bench hop_bench.cu && ./hop_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
template <bool TRY>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    if (TRY) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint32_t ld_acq(uint32_t a) { uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void st_rel(uint32_t a, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_vol(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void st_vol(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_rel(uint32_t a, uint32_t v) { asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__global__ void bench(int mode, int iters, long long* out) {
  __shared__ uint64_t bars[2];
  __shared__ uint32_t words[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), w0 = smem_u32(&words[0]), w1 = smem_u32(&words[1]);
  if (threadIdx.x == 0) {
    mbar_init(b0, 1); mbar_init(b1, (mode >= 4) ? 16 : 1);
    words[0] = words[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (mode < 4) {
    if (warp == 0) {
      for (int i = 0; i < iters; ++i) {
        if (mode <= 1) { if (lane == 0) mbar_arrive(b0); __syncwarp(); if (mode == 0) mbar_wait<true>(b1, i & 1); else mbar_wait<false>(b1, i & 1); }
        else if (mode == 2) { if (lane == 0) st_rel(w0, i + 1); __syncwarp(); while (ld_acq(w1) != (uint32_t)(i + 1)) {} }
        else { if (lane == 0) st_vol(w0, i + 1); __syncwarp(); while (ld_vol(w1) != (uint32_t)(i + 1)) {} }
      }
    } else if (warp == 1) {
      for (int i = 0; i < iters; ++i) {
        if (mode <= 1) { if (mode == 0) mbar_wait<true>(b0, i & 1); else mbar_wait<false>(b0, i & 1); if (lane == 0) mbar_arrive(b1); __syncwarp(); }
        else if (mode == 2) { while (ld_acq(w0) != (uint32_t)(i + 1)) {} if (lane == 0) st_rel(w1, i + 1); __syncwarp(); }
        else { while (ld_vol(w0) != (uint32_t)(i + 1)) {} if (lane == 0) st_vol(w1, i + 1); __syncwarp(); }
      }
    }
  } else {
    // warp 16 signals, warps 0..15 wait and answer
    if (warp == 16) {
      for (int i = 0; i < iters; ++i) {
        if (mode == 4) { if (lane == 0) mbar_arrive(b0); __syncwarp(); mbar_wait<true>(b1, i & 1); }
        else { if (lane == 0) st_rel(w0, i + 1); __syncwarp(); while (ld_acq(w1) != (uint32_t)(16 * (i + 1))) {} }
      }
    } else if (warp < 16) {
      for (int i = 0; i < iters; ++i) {
        if (mode == 4) { mbar_wait<true>(b0, i & 1); if (lane == 0) mbar_arrive(b1); __syncwarp(); }
        else { while (ld_acq(w0) != (uint32_t)(i + 1)) {} if (lane == 0) red_rel(w1, 1); __syncwarp(); }
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == (mode >= 4 ? 16 * 32 : 0)) out[0] = t1 - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 20000;
  const char* names[6] = {"mbarrier arrive + try_wait", "mbarrier arrive + test_wait spin", "st.release + ld.acquire spin", "st.volatile + ld.volatile spin",
                          "1 -> 16 warps -> 1, mbarriers (try_wait)", "1 -> 16 warps -> 1, shared-memory words"};
  for (int mode = 0; mode < 6; ++mode) {
    bench<<<1, mode >= 4 ? 17 * 32 : 64>>>(mode, iters, d);
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-44s round trip %7.1f cycles (one hop ~%6.1f)  %s\n", names[mode], (double)h / iters, (double)h / iters / 2, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
