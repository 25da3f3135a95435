// Design micro-benchmarks for the DMMA kernels (round 1): what does it cost to
// mix DFMA into a DMMA stream, and how many LDS.64 per DMMA can the SM feed?
// One CTA per SM; reports SMSP cycles per iteration next to the ideal
// (16 cycles per DMMA.8x8x4 + 2 per DFMA at the measured 37.1 TFLOP/s peak).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double& c, double a, double b) {
  asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c) : "d"(a), "d"(b));
}

// MODE 0: ND dmma then NF dfma (grouped); MODE 1: finely interleaved
template <int ND, int NF, int MODE>
__global__ void __launch_bounds__(384) k_seq(double* out, double a, double b, int iters) {
  double c0[8], c1[8], f[6];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
#pragma unroll
  for (int k = 0; k < 6; ++k) f[k] = threadIdx.x + k;
  double av[4] = {a, a + 1, a + 2, a + 3}, bv[4] = {b, b + 1, b + 2, b + 3};
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < ND; ++k) dmma(c0[k & 7], c1[k & 7], av[k & 3], bv[(k >> 2) & 3]);
#pragma unroll
      for (int k = 0; k < NF; ++k) dfma(f[k % 6], av[k & 3], bv[(k >> 1) & 3]);
    } else {
      constexpr int N = ND > NF ? ND : NF;
      int id = 0, jf = 0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        // spread the shorter stream evenly over the longer one
        if ((k + 1) * ND / N > id) { dmma(c0[id & 7], c1[id & 7], av[id & 3], bv[(id >> 2) & 3]); ++id; }
        if ((k + 1) * NF / N > jf) { dfma(f[jf % 6], av[jf & 3], bv[(jf >> 1) & 3]); ++jf; }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c0[k] + c1[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) s += f[k];
  if (s == 123.456) out[0] = s;
}

// DMMA fed by LDS.64: one B fragment load per REUSE dmma (A in registers)
template <int REUSE>
__global__ void __launch_bounds__(384) k_lds(double* out, double a, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = 1.0 / (1 + i);
  __syncthreads();
  double c0[8], c1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
  const int lane = threadIdx.x & 31;
  double av[4] = {a, a + 1, a + 2, a + 3};
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      const double b = sm[k * 32 + lane];
#pragma unroll
      for (int r = 0; r < REUSE; ++r) dmma(c0[(k * REUSE + r) & 7], c1[(k * REUSE + r) & 7], av[r & 3], b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c0[k] + c1[k];
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

static int g_sms, g_khz;
static void report(const char* name, int threads, double ms, int iters, double ideal_cycles_per_warp_iter) {
  const int warps_per_smsp_x4 = threads / 32;   // warps per SM
  const double cyc = ms * 1e-3 * g_khz * 1e3 / iters;           // SM cycles per iteration (all warps run one iteration each)
  const double ideal = ideal_cycles_per_warp_iter * warps_per_smsp_x4 / 4.0;  // SMSP cycles to serve its warps
  printf("{\"bench\": \"%s\", \"threads\": %d, \"cycles_per_iter\": %.1f, \"ideal\": %.1f, \"eff\": %.3f}\n",
         name, threads, cyc, ideal, ideal / cyc);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount; g_khz = prop.clockRate;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, g_sms, g_khz);
  double* dout; CK(cudaMalloc(&dout, 1 << 20));
  const int iters = 2000;
#define RUN_SEQ(ND, NF, MODE, TH)                                                                   \
  { double ms = time_ms([&] { k_seq<ND, NF, MODE><<<g_sms, TH>>>(dout, 1.0000001, 1e-9, iters); });  \
    report("seq_" #ND "dmma_" #NF "dfma_mode" #MODE, TH, ms, iters, ND * 16.0 + NF * 2.0); }
  for (int th : {128, 256, 384}) {
    switch (th) {
#define ALL(TH) case TH: RUN_SEQ(8, 0, 0, TH) RUN_SEQ(0, 6, 0, TH) RUN_SEQ(8, 6, 0, TH) RUN_SEQ(8, 6, 1, TH) \
      RUN_SEQ(216, 162, 0, TH) RUN_SEQ(216, 162, 1, TH) RUN_SEQ(54, 36, 0, TH) RUN_SEQ(54, 36, 1, TH) break;
      ALL(128) ALL(256) ALL(384)
    }
  }
#define RUN_LDS(R, TH)                                                                              \
  { CK(cudaFuncSetAttribute(k_lds<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 64 * 8));   \
    double ms = time_ms([&] { k_lds<R><<<g_sms, TH, 32 * 64 * 8>>>(dout, 1.0000001, iters / 4); }); \
    report("lds64_per_" #R "_dmma", TH, ms, iters / 4, 64.0 * R * 16.0); }
  RUN_LDS(1, 128) RUN_LDS(1, 256) RUN_LDS(1, 384) RUN_LDS(2, 128) RUN_LDS(2, 256) RUN_LDS(2, 384)
  return 0;
}
