// Micro-benchmarks that size the B200 rooflines the DG einsum kernels are
// measured against, and that decide the kernel design (SURVEY.md section 8(d),
// Appendix C):  FP64 vector (DFMA) peak, FP64 tensor (DMMA m8n8k4 / m16n8k16)
// peak, whether the two overlap, DFMA fed from the constant bank or from
// broadcast LDS.128, FP32 FFMA / packed FFMA2 peak, and a streaming copy.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench tools/ubench.cu
// Run  :  tools/ubench            (prints one JSON object per line)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static int g_sms = 148;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

// ---------------------------------------------------------------- DFMA -----
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
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

// ---------------------------------------------------------------- DMMA -----
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b, int iters) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) dmma884(c0[k], c1[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, double av, double bv, int iters) {
  double c[ILP][4], a[8], b[4];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = av + k;
#pragma unroll
  for (int k = 0; k < 4; ++k) b[k] = bv + k;
#pragma unroll
  for (int k = 0; k < ILP; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) dmma16816(c[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  if (s == 123.456) out[0] = s;
}

// DMMA and DFMA interleaved: NM dmma + NF dfma per inner step
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b, int iters) {
  double c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1], f[NF > 0 ? NF : 1];
#pragma unroll
  for (int k = 0; k < NM; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
#pragma unroll
  for (int k = 0; k < NF; ++k) f[k] = threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < (NM > NF ? NM : NF); ++k) {
      if (k < NM) dmma884(c0[k], c1[k], a, b);
      if (k < NF) f[k] = fma(f[k], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < NM; ++k) s += c0[k] + c1[k];
#pragma unroll
  for (int k = 0; k < NF; ++k) s += f[k];
  if (s == 123.456) out[0] = s;
}

// ------------------------------------------- DFMA with constant-bank D ------
// models "lane <-> element, operator from the constant bank": 105x35 GEMV
#define ND 3675
__constant__ double c_D[ND];
__global__ void __launch_bounds__(128) k_dfma_const(double* out, const double* u_in, int iters) {
  double u[35];
#pragma unroll
  for (int j = 0; j < 35; ++j) u[j] = u_in[(threadIdx.x + j) & 1023];
  double tot = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int rb = 0; rb < 105; rb += 5) {   // 5 rows at a time -> 5 chains
      double acc[5] = {0, 0, 0, 0, 0};
      // address is rb*35 + compile-time offset: needs LDC or indexed const operand
#pragma unroll
      for (int j = 0; j < 35; ++j) {
#pragma unroll
        for (int q = 0; q < 5; ++q) acc[q] = fma(c_D[(rb + q) * 35 + j], u[j], acc[q]);
      }
      tot += acc[0] + acc[1] + acc[2] + acc[3] + acc[4];
    }
    u[0] += tot * 1e-30;
  }
  if (tot == 123.456) out[0] = tot;
}
// fully unrolled variant: every c_D address is an immediate
template <int RB>
__device__ __forceinline__ void const_rows(const double (&u)[35], double& tot) {
  double acc[5] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 35; ++j) {
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[q] = fma(c_D[(RB + q) * 35 + j], u[j], acc[q]);
  }
  tot += acc[0] + acc[1] + acc[2] + acc[3] + acc[4];
  if constexpr (RB + 5 < 105) const_rows<RB + 5>(u, tot);
}
__global__ void __launch_bounds__(128) k_dfma_const_unrolled(double* out, const double* u_in, int iters) {
  double u[35];
#pragma unroll
  for (int j = 0; j < 35; ++j) u[j] = u_in[(threadIdx.x + j) & 1023];
  double tot = 0;
  for (int it = 0; it < iters; ++it) {
    const_rows<0>(u, tot);
    u[0] += tot * 1e-30;
  }
  if (tot == 123.456) out[0] = tot;
}

// ------------------------------------- DFMA fed by broadcast LDS.128 -------
// one LDS.128 (2 operator entries, warp-uniform) feeds 2*RE DFMAs
template <int RE>
__global__ void __launch_bounds__(128) k_dfma_lds(double* out, const double* d_in, const double* u_in, int iters) {
  __shared__ __align__(16) double sD[3680];
  for (int i = threadIdx.x; i < 3680; i += blockDim.x) sD[i] = d_in[i];
  __syncthreads();
  double u[RE][35];
#pragma unroll
  for (int e = 0; e < RE; ++e)
#pragma unroll
    for (int j = 0; j < 35; ++j) u[e][j] = u_in[(threadIdx.x + j + e) & 1023];
  double tot = 0;
  for (int it = 0; it < iters; ++it) {
    // layout: [i-pair 0..17][j 0..34][r 0..2][2]  (pairs of i adjacent -> one LDS.128)
#pragma unroll 1
    for (int ip = 0; ip < 17; ++ip) {
      double acc[RE][6];
#pragma unroll
      for (int e = 0; e < RE; ++e)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[e][q] = 0;
      const double2* p = reinterpret_cast<const double2*>(sD + ip * 210);
#pragma unroll
      for (int j = 0; j < 35; ++j) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          double2 d = p[j * 3 + r];
#pragma unroll
          for (int e = 0; e < RE; ++e) {
            acc[e][2 * r] = fma(d.x, u[e][j], acc[e][2 * r]);
            acc[e][2 * r + 1] = fma(d.y, u[e][j], acc[e][2 * r + 1]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < RE; ++e)
#pragma unroll
        for (int q = 0; q < 6; ++q) tot += acc[e][q];
    }
    u[0][0] += tot * 1e-30;
  }
  if (tot == 123.456) out[0] = tot;
}

// ---------------------------------------------------------------- FP32 -----
template <int ILP>
__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b, int iters) {
  float acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fmaf(acc[k], a, b);
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  if (s == 123.456f) out[0] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b, int iters) {
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

// ---------------------------------------------------------------- copy -----
__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = gridDim.x * (size_t)blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) out[i] = in[i];
}

// -------------------------------------------------------------- driver -----
template <class F>
static double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

static void report(const char* name, double flops_or_bytes, double ms, const char* unit, const char* note = "") {
  printf("{\"bench\": \"%s\", \"ms\": %.4f, \"rate\": %.1f, \"unit\": \"%s\", \"note\": \"%s\"}\n",
         name, ms, flops_or_bytes / (ms * 1e-3) * 1e-9, unit, note);
  fflush(stdout);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", prop.name, g_sms, prop.major, prop.minor, prop.clockRate);
  double* dout; CK(cudaMalloc(&dout, 1 << 20));
  std::vector<double> h(4096);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 / (1 + i % 17);
  double* din; CK(cudaMalloc(&din, h.size() * 8)); CK(cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(c_D, h.data(), ND * 8));

  const int iters = 4096;
  // --- FP64 vector
  for (int bps : {1, 2, 4, 8}) {
    int blocks = g_sms * bps; char note[64]; snprintf(note, 64, "256thr x %d blk/SM ILP8", bps);
    double ms = time_ms([&] { k_dfma<8><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("dfma_reg", 2.0 * 8 * iters * 256.0 * blocks, ms, "GFLOP/s", note);
  }
  {
    int blocks = g_sms * 4;
    double ms = time_ms([&] { k_dfma<16><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("dfma_reg", 2.0 * 16 * iters * 256.0 * blocks, ms, "GFLOP/s", "256thr x 4 blk/SM ILP16");
  }
  // --- FP64 tensor m8n8k4: 8*8*4*2 = 512 flop per warp instr
  for (int bps : {1, 2, 4, 8}) {
    int blocks = g_sms * bps; char note[64]; snprintf(note, 64, "256thr x %d blk/SM ILP8", bps);
    double ms = time_ms([&] { k_dmma884<8><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("dmma_m8n8k4", 512.0 * 8 * iters * 8.0 * blocks, ms, "GFLOP/s", note);
  }
  {
    int blocks = g_sms * 4;
    double ms = time_ms([&] { k_dmma884<2><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("dmma_m8n8k4", 512.0 * 2 * iters * 8.0 * blocks, ms, "GFLOP/s", "256thr x 4 blk/SM ILP2");
    ms = time_ms([&] { k_dmma884<4><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("dmma_m8n8k4", 512.0 * 4 * iters * 8.0 * blocks, ms, "GFLOP/s", "256thr x 4 blk/SM ILP4");
  }
  for (int bps : {2, 4}) {
    int blocks = g_sms * bps; char note[64]; snprintf(note, 64, "256thr x %d blk/SM ILP4", bps);
    double ms = time_ms([&] { k_dmma16816<4><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters / 4); });
    report("dmma_m16n8k16", 16.0 * 8 * 16 * 2 * 4 * (iters / 4) * 8.0 * blocks, ms, "GFLOP/s", note);
  }
  // --- do DMMA and DFMA overlap?
  {
    int blocks = g_sms * 4;
    double ms = time_ms([&] { k_mix<4, 0><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("mix_4dmma_0dfma", (512.0 * 4) * iters * 8.0 * blocks, ms, "GFLOP/s", "");
    ms = time_ms([&] { k_mix<4, 4><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("mix_4dmma_4dfma", (512.0 * 4 + 64.0 * 4) * iters * 8.0 * blocks, ms, "GFLOP/s", "if time == 4dmma_0dfma the pipes overlap");
    ms = time_ms([&] { k_mix<4, 8><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("mix_4dmma_8dfma", (512.0 * 4 + 64.0 * 8) * iters * 8.0 * blocks, ms, "GFLOP/s", "");
    ms = time_ms([&] { k_mix<2, 16><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("mix_2dmma_16dfma", (512.0 * 2 + 64.0 * 16) * iters * 8.0 * blocks, ms, "GFLOP/s", "equal flops on both");
    ms = time_ms([&] { k_mix<0, 8><<<blocks, 256>>>(dout, 1.0000001, 1e-9, iters); });
    report("mix_0dmma_8dfma", (64.0 * 8) * iters * 8.0 * blocks, ms, "GFLOP/s", "");
  }
  // --- DFMA fed from the constant bank (105x35 GEMV per thread)
  for (int bps : {2, 4}) {
    int blocks = g_sms * bps; int it2 = 64; char note[64]; snprintf(note, 64, "128thr x %d blk/SM", bps);
    double ms = time_ms([&] { k_dfma_const<<<blocks, 128>>>(dout, din, it2); });
    report("dfma_const_loop", 2.0 * ND * it2 * 128.0 * blocks, ms, "GFLOP/s", note);
    ms = time_ms([&] { k_dfma_const_unrolled<<<blocks, 128>>>(dout, din, it2); });
    report("dfma_const_unrolled", 2.0 * ND * it2 * 128.0 * blocks, ms, "GFLOP/s", note);
  }
  // --- DFMA fed from broadcast LDS.128
  for (int bps : {2, 4}) {
    int blocks = g_sms * bps; int it2 = 64; char note[64]; snprintf(note, 64, "128thr x %d blk/SM", bps);
    double ms = time_ms([&] { k_dfma_lds<1><<<blocks, 128>>>(dout, din, din, it2); });
    report("dfma_lds128_RE1", 2.0 * 17 * 210 * it2 * 128.0 * blocks, ms, "GFLOP/s", note);
  }
  for (int bps : {1, 2}) {
    int blocks = g_sms * bps; int it2 = 64; char note[64]; snprintf(note, 64, "128thr x %d blk/SM", bps);
    double ms = time_ms([&] { k_dfma_lds<2><<<blocks, 128>>>(dout, din, din, it2); });
    report("dfma_lds128_RE2", 2.0 * 2 * 17 * 210 * it2 * 128.0 * blocks, ms, "GFLOP/s", note);
  }
  // --- FP32
  for (int bps : {2, 4, 8}) {
    int blocks = g_sms * bps; char note[64]; snprintf(note, 64, "256thr x %d blk/SM ILP8", bps);
    double ms = time_ms([&] { k_ffma<8><<<blocks, 256>>>((float*)dout, 1.0000001f, 1e-9f, iters); });
    report("ffma_reg", 2.0 * 8 * iters * 256.0 * blocks, ms, "GFLOP/s", note);
    ms = time_ms([&] { k_ffma2<8><<<blocks, 256>>>((float*)dout, 1.0000001f, 1e-9f, iters); });
    report("ffma2_reg", 4.0 * 8 * iters * 256.0 * blocks, ms, "GFLOP/s", note);
  }
  // --- streaming copy, 2 GiB in + 2 GiB out
  {
    size_t n = (size_t)1 << 27;  // double2 elements = 2 GiB
    double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16)); CK(cudaMemset(a, 1, n * 16));
    for (int bps : {4, 8, 16}) {
      char note[64]; snprintf(note, 64, "k_copy %d blk/SM", bps);
      double ms = time_ms([&] { k_copy<<<g_sms * bps, 256>>>(a, b, n); });
      report("copy", 2.0 * n * 16, ms, "GB/s", note);
    }
    double ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice)); });
    report("copy", 2.0 * n * 16, ms, "GB/s", "cudaMemcpy D2D");
    cudaFree(a); cudaFree(b);
  }
  return 0;
}
