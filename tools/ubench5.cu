// DMMA.8x8x4 dependent-issue latency: NA independent accumulators used round-robin by every warp.
// cycles per DMMA (per sub-partition) = max(16 / (#warps sharing the pipe ...), latency / NA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench5 tools/ubench5.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NA>
__global__ void __launch_bounds__(512) k_chain(double* out, double a, double b, int iters, long long* cyc) {
  double c0[NA], c1[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 32; ++k) dmma(c0[k % NA], c1[k % NA], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < NA; ++k) s += c0[k] + c1[k];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NA> void run(int threads, double* out, long long* cyc) {
  const int iters = 2000;
  k_chain<NA><<<148, threads>>>(out, 1.0, 1.0, iters, cyc);
  CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const double per = (double)h / (iters * 32.0);
  printf("{\"bench\": \"dmma_chain\", \"accumulators\": %d, \"warps_per_smsp\": %.2f, \"cycles_per_dmma_per_warp\": %.1f, \"smsp_cycles_per_dmma\": %.1f}\n",
         NA, threads / 128.0, per, per / (threads / 128.0));
}
int main() {
  double* out; long long* cyc; CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8));
  for (int threads : {128, 256, 384}) {
    run<1>(threads, out, cyc); run<2>(threads, out, cyc); run<4>(threads, out, cyc);
    run<6>(threads, out, cyc); run<8>(threads, out, cyc); run<16>(threads, out, cyc);
  }
  return 0;
}
