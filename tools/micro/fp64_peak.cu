// FP64 issue-rate microbenchmark: DFMA (CUDA cores) vs mma.sync f64 (tensor cores), register-resident operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>   // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  double av[8], bv[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = b + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(av[0]), "d"(bv[0]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(av[0]), "d"(av[1]), "d"(bv[0]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(bv[0]), "d"(bv[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(av[4]), "d"(av[5]), "d"(av[6]), "d"(av[7]),
                       "d"(bv[0]), "d"(bv[1]), "d"(bv[2]), "d"(bv[3]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, (size_t)sms * 8 * 512 * 8);
  const int iters = 20000, blocks = sms * 4, threads = 512;
  float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  printf("DFMA          : %.2f TFLOP/s\n", 2.0 * blocks * threads * 16.0 * iters / (ms * 1e-3) / 1e12);
  const double mk[4] = {8 * 8 * 4, 16 * 8 * 4, 16 * 8 * 8, 16 * 8 * 16};
  ms = time_ms([&] { dmma_kernel<0><<<blocks, threads>>>(out, iters, 1.0, 1e-9); });
  printf("DMMA m8n8k4   : %.2f TFLOP/s\n", 2.0 * blocks * (threads / 32) * 8.0 * mk[0] * iters / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters, 1.0, 1e-9); });
  printf("DMMA m16n8k4  : %.2f TFLOP/s\n", 2.0 * blocks * (threads / 32) * 8.0 * mk[1] * iters / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { dmma_kernel<2><<<blocks, threads>>>(out, iters, 1.0, 1e-9); });
  printf("DMMA m16n8k8  : %.2f TFLOP/s\n", 2.0 * blocks * (threads / 32) * 8.0 * mk[2] * iters / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { dmma_kernel<3><<<blocks, threads>>>(out, iters, 1.0, 1e-9); });
  printf("DMMA m16n8k16 : %.2f TFLOP/s\n", 2.0 * blocks * (threads / 32) * 8.0 * mk[3] * iters / (ms * 1e-3) / 1e12);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
