// Phase timing of the 64 x 64 diagonal-block routine of kernels_ts.cu (clock64 stamps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../bogp_b200/csrc -o diag_time diag_time.cu
#include "../../bogp_b200/csrc/kernels_ts.cu"
#include <cstdio>
#include <vector>
namespace bogp {
void set_error(const char*, ...) {}
void count_launch(int) {}
cudaEvent_t timing_begin(cudaStream_t) { return nullptr; }
void timing_end(cudaEvent_t, cudaStream_t) {}
}
__global__ void __launch_bounds__(256) diag_time_kernel(double* S, double* Tinv, int* flag, long long* stamps) {
  extern __shared__ __align__(16) double sm[];
  double (*D)[bogp::kDiagLd] = reinterpret_cast<double (*)[bogp::kDiagLd]>(sm);
  double (*X)[bogp::kDiagLd] = D + 64;
  double* invd = reinterpret_cast<double*>(X + 64);
  double (*Tm)[16][17] = reinterpret_cast<double (*)[16][17]>(invd + 64);
  const int tid = threadIdx.x;
  for (int e = tid; e < 4096; e += 256) D[e >> 6][e & 63] = S[e];
  __syncthreads();
  bogp::diag_factor_invert(D, X, invd, Tm, tid, flag, stamps);
  for (int e = tid; e < 4096; e += 256) { S[e] = D[e >> 6][e & 63]; Tinv[e] = X[e >> 6][e & 63]; }
}
int main() {
  std::vector<double> A(4096);
  for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) A[i * 64 + j] = exp(-0.05 * (i - j) * (i - j)) + (i == j ? 0.5 : 0.0);
  double *dS, *dT; int* dflag; long long* dst;
  cudaMalloc(&dS, 4096 * 8); cudaMalloc(&dT, 4096 * 8); cudaMalloc(&dflag, 4); cudaMalloc(&dst, 32 * 8);
  cudaMemset(dflag, 0, 4);
  size_t smem = (2 * 64 * 65 + 64 + 3 * 16 * 17) * 8;
  cudaFuncSetAttribute(diag_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(dS, A.data(), 4096 * 8, cudaMemcpyHostToDevice);
    diag_time_kernel<<<1, 256, smem>>>(dS, dT, dflag, dst);
    cudaDeviceSynchronize();
  }
  long long st[32];
  cudaMemcpy(st, dst, 32 * 8, cudaMemcpyDeviceToHost);
  for (int sb = 0; sb < 4; ++sb) printf("sub-block %d: columns %lld, scaling %lld, trailing %lld\n", sb, st[5 + 4 * sb] - st[4 + 4 * sb], st[6 + 4 * sb] - st[5 + 4 * sb], (sb < 3 ? st[8 + 4 * sb] : st[1]) - st[6 + 4 * sb]);
  printf("cholesky %lld cycles, diag-inverse %lld, off-diag inverse %lld, total %lld  (%s)\n", st[1] - st[0], st[2] - st[1], st[3] - st[2], st[3] - st[0], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
