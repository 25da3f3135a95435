// Legacy tensor path (mma.sync -> HMMA) throughput on B200 for the fp32 kernel design:
// TF32 m16n8k8 and BF16 m16n8k16, fp32 accumulate, register operands, ILP 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench4 tools/ubench4.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_tf32(float* out, uint32_t a, uint32_t b, int iters) {
  float c[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3])
                   : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  if (s == 123.456f) out[0] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) k_bf16(float* out, uint32_t a, uint32_t b, int iters) {
  float c[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3])
                   : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  if (s == 123.456f) out[0] = s;
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
  float* dout; CK(cudaMalloc(&dout, 1 << 20));
  const int iters = 4096;
  for (int bps : {1, 2, 4}) {
    double ms = time_ms([&] { k_tf32<8><<<sms * bps, 256>>>(dout, 0x3f800000u, 0x3f800000u, iters); });
    printf("{\"bench\": \"mma_sync_tf32_m16n8k8\", \"blocks_per_sm\": %d, \"tflops\": %.1f, \"cycles_per_mma_per_smsp\": %.2f}\n", bps,
           2.0 * 16 * 8 * 8 * 8 * iters * 8.0 * sms * bps / (ms * 1e-3) * 1e-12,
           ms * 1e-3 * prop.clockRate * 1e3 / (8.0 * iters * 2 * bps));
    ms = time_ms([&] { k_bf16<8><<<sms * bps, 256>>>(dout, 0x3f803f80u, 0x3f803f80u, iters); });
    printf("{\"bench\": \"mma_sync_bf16_m16n8k16\", \"blocks_per_sm\": %d, \"tflops\": %.1f, \"cycles_per_mma_per_smsp\": %.2f}\n", bps,
           2.0 * 16 * 8 * 16 * 8 * iters * 8.0 * sms * bps / (ms * 1e-3) * 1e-12,
           ms * 1e-3 * prop.clockRate * 1e3 / (8.0 * iters * 2 * bps));
  }
  return 0;
}
