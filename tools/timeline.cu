// Phase timeline of the fp64 DMMA kernels at small E (no Python):  tools/timeline [E] [kind: 0 grad, 1 div, 3 lift] [threads] [fast start: 0 auto, 1 on, 2 off]
// Compiles the kernels with FNSM_TIMELINE: every warp stamps %globaltimer at entry, after the barrier set-up, after the
// operator tables are staged, when its first work item has landed, after its first item, after its last item and after
// its stores have drained.  Printed per phase: min / median / max over all warps, relative to the first CTA's entry, for
// the last of a series of back-to-back launches, plus the gap to the previous launch and the event time per launch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define FNSM_TIMELINE 1
#include "../feinsum_b200/csrc/opmat_dmma.cuh"

namespace fnsm {
std::atomic<long long> g_launches{0};
int device_info(DevInfo* out) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  out->sms = p.multiProcessorCount; out->max_smem_optin = (int)p.sharedMemPerBlockOptin;
  out->cc_major = p.major; out->cc_minor = p.minor; return 0;
}
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const long long E = argc > 1 ? atoll(argv[1]) : 100000;
  const int kind = argc > 2 ? atoi(argv[2]) : 0;
  const int threads = argc > 3 ? atoi(argv[3]) : 0;
  const int fs_mode = argc > 4 ? atoi(argv[4]) : 0;        // 0 by size, 1 fast-start instantiations, 2 round-1 kernels
  fnsm::DevInfo di; fnsm::device_info(&di);
  const int nf = kind >= 2 ? 4 : 1;
  size_t nJ, nO, nin, nout;
  if (kind == FNSM_OP_GRAD)      { nJ = 9 * E; nO = 3 * 35 * 35; nin = E * 35;     nout = 3 * E * 35; }
  else if (kind == FNSM_OP_DIV)  { nJ = 9 * E; nO = 3 * 35 * 35; nin = 3 * E * 35; nout = E * 35; }
  else                           { nJ = 4 * E; nO = 35 * 4 * 15; nin = 4 * E * 15; nout = E * 35; }
  double *dJ, *dO, *din, *dout;
  CK(cudaMalloc(&dJ, nJ * 8)); CK(cudaMalloc(&dO, nO * 8)); CK(cudaMalloc(&din, nf * nin * 8)); CK(cudaMalloc(&dout, nf * nout * 8));
  CK(cudaMemset(dJ, 0, nJ * 8)); CK(cudaMemset(dO, 0, nO * 8)); CK(cudaMemset(din, 0, nf * nin * 8));
  fnsm::OpmatRows rows{};
  for (int k = 0; k < nf; ++k) { rows.field[k] = din + k * nin; rows.out[k] = dout + k * nout; }
  fnsm_cfg cfg{}; cfg.threads = threads; cfg.reserved[2] = fs_mode << 4;
  const int n_outer = kind >= 2 ? 4 : 3, nj = kind >= 2 ? 15 : 35;
  auto launch = [&]() { return fnsm::launch_dmma(kind, dJ, dO, rows, nf, n_outer, 35, nj, E, &cfg, di, 0); };
  for (int k = 0; k < 5; ++k) if (int rc = launch()) { printf("launch failed: %d\n", rc); return 1; }
  CK(cudaDeviceSynchronize());
  const int reps = fnsm::kTlLaunches;
  unsigned zero = 0;
  CK(cudaMemcpyToSymbol(fnsm::fnsm_tl_ctr, &zero, sizeof zero));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int k = 0; k < reps; ++k) launch();
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const size_t per_launch = (size_t)160 * 16 * fnsm::kTlSlots;
  std::vector<unsigned long long> tl(per_launch * reps);
  CK(cudaMemcpyFromSymbol(tl.data(), fnsm::fnsm_tl, tl.size() * 8));
  const int nw = (threads ? threads : (kind == 0 ? 320 : kind == 1 ? 384 : 512)) / 32;
  const long long nchunks = (E + 15) / 16;
  const int grid = (int)std::min<long long>(di.sms, (nchunks * (kind >= 2 ? nf : 1) + nw - 1) / nw);
  printf("kind %d  E %lld  warps %d  grid %d  event time per launch %.2f us\n", kind, E, nw, grid, ms * 1e3 / reps);
  const char* names[9] = {"entry", "released (griddep)", "tables staged", "first item landed", "first item done",
                          "last item done", "stores drained", "raw operator landed", "... and CTA synced"};
  for (int l : {reps - 2, reps - 1}) {
    const unsigned long long* T = tl.data() + per_launch * l;
    unsigned long long t0 = ~0ull, tend = 0;
    for (int c = 0; c < grid; ++c) for (int w = 0; w < nw; ++w) {
      t0 = std::min(t0, T[(c * 16 + w) * fnsm::kTlSlots + 0]);
      tend = std::max(tend, T[(c * 16 + w) * fnsm::kTlSlots + 6]);
    }
    if (l > 0) {
      const unsigned long long* P = tl.data() + per_launch * (l - 1);
      unsigned long long pend = 0, p0 = ~0ull;
      for (int c = 0; c < grid; ++c) for (int w = 0; w < nw; ++w) {
        pend = std::max(pend, P[(c * 16 + w) * fnsm::kTlSlots + 6]);
        p0 = std::min(p0, P[(c * 16 + w) * fnsm::kTlSlots + 0]);
      }
      printf("launch %d: first entry %.2f us after the previous launch's last warp finished; previous entry-to-entry %.2f us\n",
             l, ((double)t0 - (double)pend) * 1e-3, ((double)t0 - (double)p0) * 1e-3);
    }
    printf("launch %d: in-kernel span (first entry -> last drain) %.2f us\n", l, (tend - t0) * 1e-3);
    for (int s : {0, 1, 7, 8, 2, 3, 4, 5, 6}) {
      const int slot = s >= 7 ? s + 1 : s;
      std::vector<double> v;
      for (int c = 0; c < grid; ++c) for (int w = 0; w < nw; ++w) {
        const unsigned long long* R = &T[(c * 16 + w) * fnsm::kTlSlots];
        if (R[7] == 0 && s >= 3 && s <= 4) continue;          // warp without work
        if (R[slot] == 0) continue;                              // stamp not compiled into this instantiation
        v.push_back(((double)R[slot] - (double)t0) * 1e-3);
      }
      if (v.empty()) continue;
      std::sort(v.begin(), v.end());
      printf("  %-18s min %7.2f  p10 %7.2f  median %7.2f  p90 %7.2f  max %7.2f us\n", names[s], v.front(),
             v[v.size() / 10], v[v.size() / 2], v[v.size() * 9 / 10], v.back());
    }
    std::vector<int> hist(64, 0);
    for (int c = 0; c < grid; ++c) for (int w = 0; w < nw; ++w) hist[std::min<unsigned long long>(63, T[(c * 16 + w) * fnsm::kTlSlots + 7])]++;
    printf("  items per warp:");
    for (int k = 0; k < 64; ++k) if (hist[k]) printf("  %d x %d", hist[k], k);
    printf("\n");
    // per sub-partition (warp % 4) finishing time of the slowest warp, median over CTAs
    for (int sp = 0; sp < 4; ++sp) {
      std::vector<double> v;
      for (int c = 0; c < grid; ++c) {
        double m = 0;
        for (int w = sp; w < nw; w += 4) m = std::max(m, ((double)T[(c * 16 + w) * fnsm::kTlSlots + 5] - (double)t0) * 1e-3);
        v.push_back(m);
      }
      std::sort(v.begin(), v.end());
      printf("  sub-partition %d: last item done, median over CTAs %.2f us (max %.2f)\n", sp, v[v.size() / 2], v.back());
    }
  }
  return 0;
}
