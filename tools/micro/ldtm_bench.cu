// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, alone and against a concurrent stream of
// tcgen05.mma.kind::tf32 (A in TMEM, N = 64).  Readers: R warps (quadrant = warp % 4), each looping x32 loads.
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
template <int X>
__device__ __forceinline__ float ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ float ld_sum<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]);
  return s;
}
// two x32 loads in flight before the wait
__device__ __forceinline__ float ld_sum_2x32(uint32_t taddr) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\ttcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += __uint_as_float(r[i]);
  return s;
}

// mode: 0 = readers only, 1 = MMA only, 2 = both.  wide: 0 = x32 per wait, 1 = x64 per wait
__global__ void __launch_bounds__(288, 1) bench(int readers, int mode, int wide, int iters, long long* out, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (32 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 8) {
    if (mode >= 1) {
      const uint32_t N = 64;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t db = make_desc(smem_u32(smem), 128, 256);
      long long t0 = 0;
      if (elect_one()) {
        t0 = clock64();
        for (int it = 0; it < iters; it += 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) mma_ts(tmem + 256 + 64 * (k & 1), tmem + 448 + 8 * (k & 3), db, idesc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (t0) out[0] = clock64() - t0;
    }
  } else if (warp < readers && mode != 1) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float s = 0.f;
    const long long t0 = clock64();
    const int n = iters / 4;     // reads per warp
    if (wide) for (int it = 0; it < n; it += 2) s += ld_sum_2x32(base + (uint32_t)((it * 32) & 191));
    else for (int it = 0; it < n; ++it) s += ld_sum<32>(base + (uint32_t)((it * 32) & 255));
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[1 + warp] = t1 - t0;
    sink[blockIdx.x * 288 + threadIdx.x] = s;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 16 * sizeof(long long));
  cudaMalloc(&sink, 148 * 288 * sizeof(float));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  const int iters = 8192;
  for (int wide = 0; wide < 2; ++wide)
    for (int readers : {1, 4, 8})
      for (int mode : {0, 2}) {
        cudaMemset(out, 0, 16 * sizeof(long long));
        for (int rep = 0; rep < 2; ++rep) {
          bench<<<1, 288, 32 * 1024>>>(readers, mode, wide, iters, out, sink);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[16];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < readers; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
        const double bytes = (double)readers * (iters / 4) * 32 * 32 * 4;
        printf("%s readers %d mode %d: reader cycles %lld -> %.1f B/cyc/SM (%.1f cyc per x32 load per warp)", wide ? "x64" : "x32",
               readers, mode, mx, bytes / mx, (double)mx / (iters / 4));
        if (mode) printf("; MMA %.1f cyc/MMA (ideal 32)", (double)h[0] / iters);
        printf("\n");
      }
  bench<<<1, 288, 32 * 1024>>>(0, 1, 0, iters, out, sink);
  cudaDeviceSynchronize();
  long long h0; cudaMemcpy(&h0, out, 8, cudaMemcpyDeviceToHost);
  printf("MMA alone: %.1f cyc/MMA\n", (double)h0 / iters);
  return 0;
}
