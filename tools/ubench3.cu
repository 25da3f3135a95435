// Does a DMMA stream on one warp block OTHER instruction classes issued by a
// second warp of the same SM sub-partition?  CTA = 8 warps, one CTA per SM:
// warps 0-3 (one per sub-partition) run 8-deep DMMA loops, warps 4-7 run a
// partner stream of KIND (1 IMAD, 2 LDS, 3 FFMA, 4 DFMA, 5 SHFL).  Reports the time of
// each alone and together: together ~ max -> overlap, together ~ sum -> serialised.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int KIND>
__global__ void __launch_bounds__(256) k_pair(double* out, double a, double b, int it_dmma, int it_partner) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0;
  if (warp < 4) {
    double c0[8], c1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
#pragma unroll 1
    for (int it = 0; it < it_dmma; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dmma(c0[k], c1[k], a, b);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c0[k] + c1[k];
  } else {
    if (KIND == 1) {          // integer ALU/IMAD chains
      int x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
#pragma unroll 1
      for (int it = 0; it < it_partner; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = x[k] * 3 + (x[(k + 1) & 7] ^ it);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += x[k];
    } else if (KIND == 2) {   // LDS.64
      double acc = 0; int idx = lane;
#pragma unroll 1
      for (int it = 0; it < it_partner; ++it) {
#pragma unroll
        for (int k = 0; k < 32; ++k) acc += sm[(idx + 32 * k) & 2047];
        idx = (idx + 1) & 31;
      }
      s = acc;
    } else if (KIND == 3) {   // FFMA
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = threadIdx.x + k;
#pragma unroll 1
      for (int it = 0; it < it_partner; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], (float)a, (float)b);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[k];
    } else if (KIND == 4) {   // DFMA
      double f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = threadIdx.x + k;
#pragma unroll 1
      for (int it = 0; it < it_partner; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fma(f[k], a, b);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[k];
    } else if (KIND == 5) {   // SHFL
      double v = threadIdx.x;
#pragma unroll 1
      for (int it = 0; it < it_partner; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v += __shfl_xor_sync(0xffffffffu, v, 1 + (k & 3));
      }
      s = v;
    }
  }
  if (s == 123.456) out[0] = s;
}

template <class F> static double time_ms(F f) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 1 << 20));
  const int itd = 4000;
#define RUN(KIND, NAME, ITP)                                                                        \
  { double td = time_ms([&] { k_pair<KIND><<<sms, 256>>>(dout, 1.0000001, 1e-9, itd, 0); });        \
    double tp = time_ms([&] { k_pair<KIND><<<sms, 256>>>(dout, 1.0000001, 1e-9, 0, ITP); });        \
    double tb = time_ms([&] { k_pair<KIND><<<sms, 256>>>(dout, 1.0000001, 1e-9, itd, ITP); });      \
    printf("{\"partner\": \"%s\", \"dmma_alone_ms\": %.4f, \"partner_alone_ms\": %.4f, \"together_ms\": %.4f, " \
           "\"serialisation\": %.3f}\n", NAME, td, tp, tb, (tb - (td > tp ? td : tp)) / (td < tp ? td : tp)); }
  RUN(1, "imad", 16000) RUN(2, "lds64", 10000) RUN(3, "ffma", 8000) RUN(4, "dfma", 4000) RUN(5, "shfl", 8000)
  return 0;
}
