// Device peak measurements for the roofline model (replaces the static table
// of the reference, src/feinsum/data/device_info.py:4-28, by numbers measured
// on the running board; SURVEY.md section 8(d) asks for FP64/FP32 peaks
// because MEASURED_PEAKS.json only carries HBM and bf16).
#include "common.cuh"

namespace fnsm {

template <int ILP>
__global__ void __launch_bounds__(256) k_peak_dfma(double* out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_peak_ffma2(float* out, float a, float b, int iters) {
  // packed FP32: two FMAs per lane per instruction (FFMA2 on sm_100)
  unsigned long long acc[ILP], av, bv;
  asm("mov.b64 %0, {%1,%1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1,%1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int k = 0; k < ILP; ++k) { float x = threadIdx.x + k; asm("mov.b64 %0, {%1,%1};" : "=l"(acc[k]) : "f"(x)); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(av), "l"(bv));
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k])); s += lo + hi; }
  if (s == 123.456f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_peak_dmma(double* out, double a, double b, int iters) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_peak_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = gridDim.x * (size_t)blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = __ldcs(in + i), b = __ldcs(in + i + stride), c = __ldcs(in + i + 2 * stride), d = __ldcs(in + i + 3 * stride);
    __stcs(out + i, a); __stcs(out + i + stride, b); __stcs(out + i + 2 * stride, c); __stcs(out + i + 3 * stride, d);
  }
  for (; i < n; i += stride) __stcs(out + i, __ldcs(in + i));
}

// sustained = true: keep the pipe busy for ~0.7 s first, so the figure is the one the board
// holds under its power cap (B200: FP64 tensor work pulls the 1000 W cap after ~0.3 s and the SM
// clock settles below the 1965 MHz a short burst sees), then average instead of taking the best
template <class F>
static int best_ms(F launch, double* ms_out, bool sustained = false) {
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return (int)cudaGetLastError();
  launch(); launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  if (sustained) {
    float elapsed = 0;
    cudaEventRecord(e0);
    while (elapsed < 700.f) {
      for (int r = 0; r < 8; ++r) launch();
      cudaEventRecord(e1);
      if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return (int)e;
      cudaEventElapsedTime(&elapsed, e0, e1);
    }
    const int reps = 16;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return (int)e;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_out = ms / reps;
    e = cudaGetLastError();
    return e == cudaSuccess ? FNSM_OK : (int)e;
  }
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return (int)e;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_out = best;
  e = cudaGetLastError();
  return e == cudaSuccess ? FNSM_OK : (int)e;
}

}  // namespace fnsm

extern "C" int fnsm_b200_measure_peak(int32_t which, double* result) {
  using namespace fnsm;
  const bool sustained = (which & 16) != 0;     // + 16: sustained (power-capped) instead of burst
  which &= ~16;
  if (!result || which < 0 || which > 3) return FNSM_E_BAD_ARG;
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  double* scratch = nullptr;
  cudaError_t e = cudaMalloc(&scratch, 4096);
  if (e != cudaSuccess) return (int)e;
  int rc = FNSM_OK;
  double ms = 0;
  const int iters = 8192, blocks = di.sms * 4;
  if (which == 0) {
    rc = best_ms([&] { k_peak_dfma<8><<<blocks, 256>>>(scratch, 1.0000001, 1e-9, iters); g_launches++; }, &ms, sustained);
    *result = 2.0 * 8 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-9;
  } else if (which == 1) {
    rc = best_ms([&] { k_peak_ffma2<8><<<blocks, 256>>>((float*)scratch, 1.0000001f, 1e-9f, iters); g_launches++; }, &ms, sustained);
    *result = 4.0 * 8 * iters * 256.0 * blocks / (ms * 1e-3) * 1e-9;
  } else if (which == 3) {
    rc = best_ms([&] { k_peak_dmma<8><<<blocks, 256>>>(scratch, 1.0000001, 1e-9, iters); g_launches++; }, &ms, sustained);
    *result = 512.0 * 8 * iters * 8.0 * blocks / (ms * 1e-3) * 1e-9;
  } else {
    const size_t n = (size_t)1 << 26;  // 1 GiB in + 1 GiB out
    double2 *a = nullptr, *b = nullptr;
    if ((e = cudaMalloc(&a, n * 16)) != cudaSuccess || (e = cudaMalloc(&b, n * 16)) != cudaSuccess) {
      cudaFree(a); cudaFree(scratch); return (int)e;
    }
    cudaMemset(a, 1, n * 16);
    rc = best_ms([&] { k_peak_copy<<<di.sms * 8, 256>>>(a, b, n); g_launches++; }, &ms, sustained);
    *result = 2.0 * n * 16 / (ms * 1e-3) * 1e-9;
    cudaFree(a); cudaFree(b);
  }
  cudaFree(scratch);
  return rc;
}
