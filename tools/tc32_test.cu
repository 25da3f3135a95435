// Standalone check + timing of the tcgen05 fp32 kernels (no Python): tools/tc32_test [E] [reps]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <random>
#include "../feinsum_b200/csrc/opmat_tc32.cuh"

namespace fnsm {
std::atomic<long long> g_launches{0};
int device_info(DevInfo* out) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  out->sms = p.multiProcessorCount; out->max_smem_optin = (int)p.sharedMemPerBlockOptin;
  out->cc_major = p.major; out->cc_minor = p.minor; return 0;
}
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static int test_lift(bool fe, long long E, int reps, const fnsm::DevInfo& di) {
  std::mt19937 rng(2);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  const int nf = 4;
  std::vector<float> J(4 * E), O(35 * 4 * 15), v((size_t)nf * 4 * E * 15), out((size_t)nf * E * 35);
  for (auto& x : J) x = U(rng);
  for (auto& x : O) x = U(rng);
  for (auto& x : v) x = U(rng);
  float *dJ, *dO, *dv, *dout;
  CK(cudaMalloc(&dJ, J.size() * 4)); CK(cudaMalloc(&dO, O.size() * 4)); CK(cudaMalloc(&dv, v.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4));
  CK(cudaMemcpy(dJ, J.data(), J.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dO, O.data(), O.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, out.size() * 4));
  fnsm::OpmatRows rows{};
  for (int k = 0; k < nf; ++k) { rows.field[k] = dv + (size_t)k * 4 * E * 15; rows.out[k] = dout + (size_t)k * E * 35; }
  const int kind = fe ? FNSM_OP_LIFT_FE : FNSM_OP_LIFT_EF;
  int rc = fnsm::launch_lift_tc32<35, 15>(kind, dJ, dO, rows, nf, E, di, 0, false);
  printf("launch rc=%d\n", rc);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0; long long bad = 0, checked = 0;
  const long long step = E > 20000 ? E / 20000 : 1;
  for (int k = 0; k < nf; ++k)
    for (long long e = 0; e < E; e += (e + step < E - 300 ? step : 1))
      for (int i = 0; i < 35; ++i) {
        double ref = 0;
        for (int f = 0; f < 4; ++f) {
          const double jf = fe ? J[(size_t)f * E + e] : J[(size_t)e * 4 + f];
          for (int j = 0; j < 15; ++j) {
            const double op = fe ? O[(i * 4 + f) * 15 + j] : O[(f * 35 + i) * 15 + j];
            ref += op * jf * v[(((size_t)k * 4 + f) * E + e) * 15 + j];
          }
        }
        const double got = out[((size_t)k * E + e) * 35 + i];
        const double err = fabs(got - ref) / fmax(fabs(ref), 1e-30);
        if (!(err < 1e-5)) { if (bad < 10) printf("bad k=%d e=%lld i=%d got=%g ref=%g\n", k, e, i, got, ref); ++bad; }
        if (err > worst || err != err) worst = err;
        ++checked;
      }
  printf("lift_%s E=%lld checked=%lld bad=%lld worst_rel=%.3e\n", fe ? "fe" : "ef", E, checked, bad, worst);
  if (reps > 0) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int k = 0; k < 3; ++k) fnsm::launch_lift_tc32<35, 15>(kind, dJ, dO, rows, nf, E, di, 0, false);
    cudaEventRecord(a);
    for (int k = 0; k < reps; ++k) fnsm::launch_lift_tc32<35, 15>(kind, dJ, dO, rows, nf, E, di, 0, false);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    printf("lift_tc32 E=%lld: %.4f ms  %.1f GB/s  %.1f TFLOP/s(17040/elt)\n", E, ms, 1536.0 * E / ms * 1e-6, 17040.0 * E / ms * 1e-9);
  }
  cudaFree(dJ); cudaFree(dO); cudaFree(dv); cudaFree(dout);
  return bad ? 2 : 0;
}

static int test_div(long long E, int reps, const fnsm::DevInfo& di) {
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  std::vector<float> J(9 * E), D(3 * 35 * 35), u((size_t)3 * E * 35), out((size_t)E * 35);
  for (auto& v : J) v = U(rng);
  for (auto& v : D) v = U(rng);
  for (auto& v : u) v = U(rng);
  float *dJ, *dD, *du, *dout;
  CK(cudaMalloc(&dJ, J.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&du, u.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4));
  CK(cudaMemcpy(dJ, J.data(), J.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(du, u.data(), u.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, out.size() * 4));
  int rc = fnsm::launch_div_tc32<35>(dJ, dD, du, dout, E, di, 0, false);
  printf("launch rc=%d\n", rc);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0; long long bad = 0, checked = 0;
  const long long step = E > 20000 ? E / 20000 : 1;
  for (long long e = 0; e < E; e += (e + step < E - 300 ? step : 1))
    for (int i = 0; i < 35; ++i) {
      double ref = 0;
      for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 35; ++j) {
          double w = 0;
          for (int x = 0; x < 3; ++x) w += (double)J[(size_t)(3 * x + r) * E + e] * u[((size_t)x * E + e) * 35 + j];
          ref += (double)D[(r * 35 + i) * 35 + j] * w;
        }
      const double got = out[(size_t)e * 35 + i];
      const double err = fabs(got - ref) / fmax(fabs(ref), 1e-30);
      if (!(err < 1e-5)) { if (bad < 10) printf("bad e=%lld i=%d got=%g ref=%g\n", e, i, got, ref); ++bad; }
      if (err > worst || err != err) worst = err;
      ++checked;
    }
  printf("div E=%lld checked=%lld bad=%lld worst_rel=%.3e\n", E, checked, bad, worst);
  if (reps > 0) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int k = 0; k < 3; ++k) fnsm::launch_div_tc32<35>(dJ, dD, du, dout, E, di, 0, false);
    cudaEventRecord(a);
    for (int k = 0; k < reps; ++k) fnsm::launch_div_tc32<35>(dJ, dD, du, dout, E, di, 0, false);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    printf("div_tc32 E=%lld: %.4f ms  %.1f GB/s  %.1f TFLOP/s(7980/elt)\n", E, ms, 596.0 * E / ms * 1e-6, 7980.0 * E / ms * 1e-9);
  }
  cudaFree(dJ); cudaFree(dD); cudaFree(du); cudaFree(dout);
  return bad ? 2 : 0;
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "grad";
  long long E = argc > 2 ? atoll(argv[2]) : 40028;
  int reps = argc > 3 ? atoi(argv[3]) : 0;
  fnsm::DevInfo di; fnsm::device_info(&di);
  if (!strcmp(mode, "lift_fe")) return test_lift(true, E, reps, di);
  if (!strcmp(mode, "lift_ef")) return test_lift(false, E, reps, di);
  if (!strcmp(mode, "div")) return test_div(E, reps, di);
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  std::vector<float> J(9 * E), D(3 * 35 * 35), u(E * 35), out(3 * E * 35);
  for (auto& v : J) v = U(rng);
  for (auto& v : D) v = U(rng);
  for (auto& v : u) v = U(rng);
  float *dJ, *dD, *du, *dout;
  CK(cudaMalloc(&dJ, J.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&du, u.size() * 4)); CK(cudaMalloc(&dout, out.size() * 4));
  CK(cudaMemcpy(dJ, J.data(), J.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(du, u.data(), u.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, out.size() * 4));
  int rc = fnsm::launch_grad_tc32<35>(dJ, dD, du, dout, E, di, 0, false);
  printf("launch rc=%d\n", rc);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0; long long bad = 0, checked = 0;
  const long long step = E > 20000 ? E / 20000 : 1;
  for (long long e = 0; e < E; e += (e + step < E - 300 ? step : 1)) {
    for (int i = 0; i < 35; ++i) {
      double T[3] = {0, 0, 0};
      for (int r = 0; r < 3; ++r) for (int j = 0; j < 35; ++j) T[r] += (double)D[(r * 35 + i) * 35 + j] * u[e * 35 + j];
      for (int x = 0; x < 3; ++x) {
        double ref = 0; for (int r = 0; r < 3; ++r) ref += (double)J[(3 * x + r) * E + e] * T[r];
        double got = out[((long long)x * E + e) * 35 + i];
        double err = fabs(got - ref) / fmax(fabs(ref), 1e-30);
        if (!(err < 1e-5)) { if (bad < 10) printf("bad e=%lld x=%d i=%d got=%g ref=%g\n", e, x, i, got, ref); ++bad; }
        if (err > worst || err != err) worst = err;
        ++checked;
      }
    }
  }
  printf("E=%lld checked=%lld bad=%lld worst_rel=%.3e\n", E, checked, bad, worst);
  if (reps > 0) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int k = 0; k < 3; ++k) fnsm::launch_grad_tc32<35>(dJ, dD, du, dout, E, di, 0, false);
    cudaEventRecord(a);
    for (int k = 0; k < reps; ++k) fnsm::launch_grad_tc32<35>(dJ, dD, du, dout, E, di, 0, false);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
    printf("grad_tc32 E=%lld: %.4f ms  %.1f GB/s  %.1f TFLOP/s(7980/elt)\n", E, ms, 596.0 * E / ms * 1e-6, 7980.0 * E / ms * 1e-9);
  }
  return bad ? 2 : 0;
}
