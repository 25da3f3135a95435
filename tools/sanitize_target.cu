// Small-E launches of every kernel family through the C ABI, for compute-sanitizer
// (tests/test_gpu_sanitizer.py runs this binary under memcheck / racecheck / synccheck / initcheck).
// Pure C-ABI client: no torch, no Python -- also the minimal example of binding include/fnsm_b200.h from C++.
//
//   sanitize_target [family ...]      families: dmma dmma_plain dmma_gen tc32 tf32 tf32_gen simt tp generic wave hexd se
//
// Every opmat result is cross-checked against the simt variant of the same einsum (max relative
// difference printed; non-zero exit on mismatch), so a sanitizer-clean run is also a run that computed.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "fnsm_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); std::exit(2); } } while (0)
#define FN(x) do { int rc_ = (x); if (rc_ != 0) { std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, fnsm_b200_strerror(rc_)); std::exit(3); } } while (0)

static unsigned long long g_seed = 0x9E3779B97F4A7C15ull;
static double rnd() { g_seed = g_seed * 6364136223846793005ull + 1442695040888963407ull; return (double)(g_seed >> 11) * (1.0 / 9007199254740992.0); }

// Device buffer with canary zones in front of and behind the payload (compute-sanitizer is closed on this
// pool, so out-of-bounds WRITES are caught here: every guard byte must still hold the canary afterwards) and an
// optional offset that makes the base not 16-byte aligned.
static int g_guard_fail = 0;
template <class T> struct Buf {
  static constexpr size_t G = 1024;          // guard elements on each side
  T* d = nullptr; T* raw = nullptr; size_t n = 0; size_t off = 0;
  Buf(size_t n_, bool random, size_t off_ = 0) : n(n_), off(off_) {
    CK(cudaMalloc(&raw, (n + off + 2 * G) * sizeof(T)));
    CK(cudaMemset(raw, 0xA5, (n + off + 2 * G) * sizeof(T)));
    d = raw + G + off;
    std::vector<T> h(n);
    for (auto& v : h) v = random ? (T)rnd() : (T)0;
    CK(cudaMemcpy(d, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
  }
  ~Buf() {
    std::vector<unsigned char> g((G + off) * sizeof(T)), t(G * sizeof(T));
    CK(cudaMemcpy(g.data(), raw, g.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t.data(), d + n, t.size(), cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (unsigned char c : g) bad += c != 0xA5;
    for (unsigned char c : t) bad += c != 0xA5;
    if (bad) { std::fprintf(stderr, "GUARD VIOLATION: %zu canary bytes overwritten around a %zu-element buffer\n", bad, n); g_guard_fail = 1; }
    cudaFree(raw);
  }
  std::vector<T> host() const { std::vector<T> h(n); CK(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost)); return h; }
};

template <class T> static double max_rel(const std::vector<T>& a, const std::vector<T>& b) {
  double m = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) { s = std::fmax(s, std::fabs((double)b[i])); }
  for (size_t i = 0; i < a.size(); ++i) m = std::fmax(m, std::fabs((double)a[i] - (double)b[i]));
  return s > 0 ? m / s : m;
}

static int g_fail = 0;

// one opmat einsum with launch variant `variant` vs the simt variant
template <class T>
static void opmat(const char* what, int kind, int nd, int nfd, long long E, int variant, int b, size_t misalign, int threads = 0) {
  const bool lift = kind >= FNSM_OP_LIFT_EF;
  const int n_outer = lift ? 4 : 3, ni = nd, nj = lift ? nfd : nd;
  Buf<T> jac((size_t)(lift ? 4 : 9) * E, true, misalign), op((size_t)n_outer * ni * nj, true);
  const size_t fsz = (size_t)(kind == FNSM_OP_GRAD ? 1 : n_outer) * E * nj;
  const size_t osz = (size_t)(kind == FNSM_OP_GRAD ? 3 : 1) * E * ni;
  std::vector<Buf<T>*> f, o, r;
  std::vector<const void*> fp; std::vector<void*> opn, rp;
  for (int k = 0; k < b; ++k) {
    f.push_back(new Buf<T>(fsz, true, misalign)); o.push_back(new Buf<T>(osz, false, misalign)); r.push_back(new Buf<T>(osz, false));
    fp.push_back(f[k]->d); opn.push_back(o[k]->d); rp.push_back(r[k]->d);
  }
  const int dt = sizeof(T) == 8 ? FNSM_F64 : FNSM_F32;
  fnsm_cfg cfg; std::memset(&cfg, 0, sizeof cfg); cfg.variant = variant; cfg.threads = threads;
  FN(fnsm_b200_opmat_batch(kind, dt, jac.d, op.d, fp.data(), opn.data(), b, n_outer, ni, nj, E, &cfg, nullptr));
  fnsm_cfg ref; std::memset(&ref, 0, sizeof ref); ref.variant = 2;
  FN(fnsm_b200_opmat_batch(kind, dt, jac.d, op.d, fp.data(), rp.data(), b, n_outer, ni, nj, E, &ref, nullptr));
  CK(cudaDeviceSynchronize());
  double worst = 0;
  for (int k = 0; k < b; ++k) worst = std::fmax(worst, max_rel(o[k]->host(), r[k]->host()));
  const double tol = sizeof(T) == 8 ? 1e-13 : 1e-5;
  std::printf("%-44s E=%-7lld variant=%d  max rel diff vs simt %.3g %s\n", what, E, variant, worst, worst < tol ? "ok" : "MISMATCH");
  if (!(worst < tol)) g_fail = 1;
  for (int k = 0; k < b; ++k) { delete f[k]; delete o[k]; delete r[k]; }
}

static void tensor_product(long long E) {
  Buf<double> A((size_t)E * 512, true), M(64, true), out((size_t)E * 512, false);
  for (int mode = 0; mode < 3; ++mode) FN(fnsm_b200_tensor_product(FNSM_F64, A.d, M.d, out.d, 8, mode, E, nullptr, nullptr));
  Buf<float> Af((size_t)E * 343, true), Mf(49, true), outf((size_t)E * 343, false);
  for (int mode = 0; mode < 3; ++mode) FN(fnsm_b200_tensor_product(FNSM_F32, Af.d, Mf.d, outf.d, 7, mode, E, nullptr, nullptr));
  CK(cudaDeviceSynchronize());
  std::printf("%-44s E=%-7lld ok\n", "tensor-product n1d=8 fp64, n1d=7 fp32, 3 modes", E);
}

static void generic(long long E) {
  // out[e,i] = sum_j A[e,i,j] x[e,j]
  fnsm_einsum_desc d; std::memset(&d, 0, sizeof d);
  d.n_free = 2; d.n_sum = 1; d.n_operands = 2; d.dtype = FNSM_F64;
  d.extent[0] = E; d.extent[1] = 5; d.extent[2] = 7;
  d.out_stride[0] = 5; d.out_stride[1] = 1;
  d.in_stride[0][0] = 35; d.in_stride[0][1] = 7; d.in_stride[0][2] = 1;
  d.in_stride[1][0] = 7; d.in_stride[1][2] = 1;
  Buf<double> A((size_t)E * 35, true), x((size_t)E * 7, true), out((size_t)E * 5, false);
  const void* in[2] = {A.d, x.d}; void* o[1] = {out.d};
  FN(fnsm_b200_generic_einsum(&d, 1, in, o, nullptr));
  CK(cudaDeviceSynchronize());
  std::printf("%-44s E=%-7lld ok\n", "generic eij,ej->ei", E);
}

template <class T>
static void wave_t(long long E, int dtype) {
  Buf<T> J(9 * E, true), D(3 * 35 * 35, true), v(3 * E * 35, true), u(E * 35, true), L(35 * 60, true), Jf(4 * E, true);
  Buf<T> dv(E * 35, false), gr(3 * E * 35, false);
  std::vector<Buf<T>*> F, lo;
  fnsm_wave_args a; std::memset(&a, 0, sizeof a);
  a.J = J.d; a.D = D.d; a.v = v.d; a.u = u.d; a.L = L.d; a.Jface = Jf.d; a.div_out = dv.d; a.grad_out = gr.d;
  for (int k = 0; k < 4; ++k) { F.push_back(new Buf<T>(4 * E * 15, true)); lo.push_back(new Buf<T>(E * 35, false)); a.F[k] = F[k]->d; a.lift_out[k] = lo[k]->d; }
  FN(fnsm_b200_wave3d_fused(dtype, &a, E, nullptr, nullptr));
  // reference: the three einsums one by one through the simt kernels
  fnsm_cfg ref; std::memset(&ref, 0, sizeof ref); ref.variant = 2;
  Buf<T> dv2(E * 35, false), gr2(3 * E * 35, false);
  std::vector<Buf<T>*> lo2; std::vector<const void*> fp; std::vector<void*> op2;
  for (int k = 0; k < 4; ++k) { lo2.push_back(new Buf<T>(E * 35, false)); fp.push_back(F[k]->d); op2.push_back(lo2[k]->d); }
  const void* f1[1] = {v.d}; void* o1[1] = {dv2.d};
  FN(fnsm_b200_opmat_batch(FNSM_OP_DIV, dtype, J.d, D.d, f1, o1, 1, 3, 35, 35, E, &ref, nullptr));
  const void* f2[1] = {u.d}; void* o2[1] = {gr2.d};
  FN(fnsm_b200_opmat_batch(FNSM_OP_GRAD, dtype, J.d, D.d, f2, o2, 1, 3, 35, 35, E, &ref, nullptr));
  FN(fnsm_b200_opmat_batch(FNSM_OP_LIFT_FE, dtype, Jf.d, L.d, fp.data(), op2.data(), 4, 4, 35, 15, E, &ref, nullptr));
  CK(cudaDeviceSynchronize());
  double worst = std::fmax(max_rel(dv.host(), dv2.host()), max_rel(gr.host(), gr2.host()));
  for (int k = 0; k < 4; ++k) worst = std::fmax(worst, max_rel(lo[k]->host(), lo2[k]->host()));
  const double tol = sizeof(T) == 8 ? 1e-13 : 1e-5;
  std::printf("%-44s E=%-7lld            max rel diff vs simt %.3g %s\n", sizeof(T) == 8 ? "wave_3d_p4 fp64" : "wave_3d_p4 fp32", E, worst, worst < tol ? "ok" : "MISMATCH");
  if (!(worst < tol)) g_fail = 1;
  for (int k = 0; k < 4; ++k) { delete F[k]; delete lo[k]; delete lo2[k]; }
}
static void wave(long long E, int dtype) { if (dtype == FNSM_F64) wave_t<double>(E, dtype); else wave_t<float>(E, dtype); }

int main(int argc, char** argv) {
  std::vector<std::string> fams;
  for (int i = 1; i < argc; ++i) fams.push_back(argv[i]);
  if (fams.empty()) fams = {"dmma", "dmma_plain", "dmma_gen", "tc32", "tf32", "tf32_gen", "simt", "tp", "generic", "wave"};
  for (const auto& f : fams) {
    if (f == "dmma") {            // fp64 p = 4, TMA path (E even, aligned), every warp count in use
      opmat<double>("dmma grad p4 (TMA)", FNSM_OP_GRAD, 35, 15, 6000, 1, 1, 0);
      opmat<double>("dmma div p4 (TMA, direct stores)", FNSM_OP_DIV, 35, 15, 6000, 1, 1, 0);
      opmat<double>("dmma div p4 (TMA, staged, 10 warps)", FNSM_OP_DIV, 35, 15, 6000, 1, 1, 0, 320);
      opmat<double>("dmma lift_fe p4 b=4 (TMA)", FNSM_OP_LIFT_FE, 35, 15, 3000, 1, 4, 0);
      opmat<double>("dmma lift_ef p4 b=3 (TMA)", FNSM_OP_LIFT_EF, 35, 15, 3000, 1, 3, 0);
    } else if (f == "dmma_plain") {   // odd E / misaligned bases: plain-load producer
      opmat<double>("dmma grad p4 (odd E)", FNSM_OP_GRAD, 35, 15, 5999, 1, 1, 0);
      opmat<double>("dmma div p4 (misaligned)", FNSM_OP_DIV, 35, 15, 6000, 1, 1, 1);
      opmat<double>("dmma lift_fe p4 b=4 (odd E)", FNSM_OP_LIFT_FE, 35, 15, 2999, 1, 4, 0);
    } else if (f == "dmma_gen") {
      opmat<double>("dmma_gen grad p2", FNSM_OP_GRAD, 10, 6, 9001, 1, 1, 0);
      opmat<double>("dmma_gen div p3", FNSM_OP_DIV, 20, 10, 9001, 1, 1, 0);
      opmat<double>("dmma_gen lift_fe p1 b=4", FNSM_OP_LIFT_FE, 4, 3, 9001, 1, 4, 0);
    } else if (f == "tc32") {      // tcgen05 + TMEM + TMA
      opmat<float>("tc32 grad p4", FNSM_OP_GRAD, 35, 15, 8000, 3, 1, 0);
      opmat<float>("tc32 div p4", FNSM_OP_DIV, 35, 15, 8000, 3, 1, 0);
      opmat<float>("tc32 lift_fe p4 b=4", FNSM_OP_LIFT_FE, 35, 15, 4000, 3, 4, 0);
      opmat<float>("tc32 lift_ef p2 b=2", FNSM_OP_LIFT_EF, 10, 6, 8000, 3, 2, 0);
      opmat<float>("tc32 grad p1", FNSM_OP_GRAD, 4, 3, 8000, 3, 1, 0);
    } else if (f == "tf32") {      // fp32 operands that do not qualify for TMA (auto picks the non-TMA producer)
      opmat<float>("fp32 grad p4, E % 4 != 0 (auto)", FNSM_OP_GRAD, 35, 15, 6001, 0, 1, 0);
      opmat<float>("fp32 div p4, misaligned (auto)", FNSM_OP_DIV, 35, 15, 6000, 0, 1, 1);
      opmat<float>("fp32 lift_fe p4 b=4, E % 4 != 0 (auto)", FNSM_OP_LIFT_FE, 35, 15, 3002, 0, 4, 0);
      opmat<float>("tf32 mma.sync grad p4", FNSM_OP_GRAD, 35, 15, 6000, 1, 1, 0);
      opmat<float>("tf32 mma.sync div p4", FNSM_OP_DIV, 35, 15, 6000, 1, 1, 0);
      opmat<float>("tf32 mma.sync lift_fe p4", FNSM_OP_LIFT_FE, 35, 15, 3000, 1, 4, 0);
    } else if (f == "tf32_gen") {
      opmat<float>("tf32_gen grad p3", FNSM_OP_GRAD, 20, 10, 9001, 1, 1, 0);
      opmat<float>("tf32_gen lift_fe p2 b=4", FNSM_OP_LIFT_FE, 10, 6, 9001, 1, 4, 0);
    } else if (f == "simt") {
      opmat<double>("simt grad 15 dofs", FNSM_OP_GRAD, 15, 5, 3001, 2, 1, 0);
      opmat<float>("simt div p5 (56 dofs)", FNSM_OP_DIV, 56, 21, 1001, 2, 1, 0);
    } else if (f == "tp") {
      tensor_product(3001);
    } else if (f == "generic") {
      generic(3001);
    } else if (f == "wave") {
      wave(6000, FNSM_F64);
      wave(8000, FNSM_F32);
    } else {
      std::fprintf(stderr, "unknown family %s\n", f.c_str());
      return 4;
    }
  }
  std::printf("launches through the ABI: %lld\n", (long long)fnsm_b200_launch_count());
  if (g_guard_fail) std::printf("GUARD VIOLATION\n");
  return g_fail | (g_guard_fail << 1);
}
