// Tensor-core peak on this GPU as the Bartlett kernel can use it: every SM issues tcgen05.mma (M = 128, A from TMEM
// or shared memory, B from shared memory) back to back for a few milliseconds; rate = flop / CUDA-event time, i.e. at
// the clock a short burst really runs at.  kind::tf32 (K = 8) and, for reference, kind::f16 (K = 16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_peak tc_peak.cu && ./tc_peak   -> one JSON object per line
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
template <int KIND, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_s, uint64_t b, uint32_t idesc) {
  if (KIND == 0) {
    if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_s), "l"(b), "r"(idesc) : "memory");
  } else {
    if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_s), "l"(b), "r"(idesc) : "memory");
  }
}

template <int KIND, bool TS>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    // instruction descriptor: D fp32 (bit 4), A/B format at bits 7 / 10 (tf32 = 2, f16 = 0), N >> 3 at 17, M >> 4 at 24
    const uint32_t fmt = KIND == 0 ? 2u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sA = smem_u32(smem), sB = sA + 32 * 1024;
    const uint64_t da = make_desc(sA, 128, 256), db = make_desc(sB, 128, 256);
    const uint32_t d0 = tmem, d1 = tmem + 256, a0 = tmem + 496;
    if (elect_one()) {
      for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mma<KIND, TS>((k & 1) ? d1 : d0, a0, da, db, idesc);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int KIND, bool TS>
static int run(int N, int sms, double target_ms) {
  cudaFuncSetAttribute(bench<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int K = KIND == 0 ? 8 : 16;
  int iters = 1 << 14;
  float best = 0.f, ms = 0.f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    bench<KIND, TS><<<sms, 128, 64 * 1024>>>(N, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 0) { iters = (int)(iters * target_ms / ms) & ~7; continue; }      // size the run from the first probe
    const double tf = 2.0 * 128 * N * K * (double)iters * sms / (ms * 1e-3) * 1e-12;
    if (tf > best) best = (float)tf;
  }
  printf("{\"kind\": \"%s\", \"a_operand\": \"%s\", \"N\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"cycles_per_mma_at_1965MHz\": %.2f}\n",
         KIND == 0 ? "tf32" : "f16", TS ? "tmem" : "smem", N, ms, best, 2.0 * 128 * N * K * sms / (best * 1e12) * 1.965e9);
  return 0;
}

int main(int argc, char** argv) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const double ms = argc > 1 ? atof(argv[1]) : 5.0;       // the Bartlett kernel runs ~0.1-7 ms bursts
  for (int N : {64, 128, 256}) { if (run<0, true>(N, sms, ms)) return 1; }
  for (int N : {64, 128, 256}) { if (run<0, false>(N, sms, ms)) return 1; }
  for (int N : {64, 256}) { if (run<1, true>(N, sms, ms)) return 1; }
  if (run<0, true>(256, sms, 2000.0)) return 1;           // sustained: two seconds of back-to-back MMAs
  return 0;
}
