// Fused three-mode hex derivative: the three tensor-product applications
//     d0 = eabc,ia->eibc     d1 = eabc,ib->eaic     d2 = eabc,ic->eabi
// of 1-D operators M0, M1, M2 (8 x 8, p = 7 hexes) to the SAME field A(E,8,8,8), with A read ONCE:
// 4 KB in + 12 KB out = 16 KB per element instead of 3 x 8 KB for three fnsm_b200_tensor_product calls.
// HBM bound (1.5 flop/B).
//
// One warp owns one element at a time:
//   * the element's 512 doubles travel global -> shared with 16-byte cp.async (8 per lane) into a warp-private
//     ring of NBUF buffers, NBUF - 1 elements ahead of the one being computed -- no CTA-wide barrier anywhere;
//     planes A[a][.][.] sit 72 doubles apart (8 doubles of padding) so that the mode-1 reads, which stride over b,
//     are bank-conflict free;
//   * mode 0: lane <-> (b,c)-pair: 8 LDS.128 down the a axis, 128 DFMA against M0 (broadcast LDS), 8 stores of
//     512 contiguous bytes per warp;   mode 1: lane <-> (a, c-pair): 8 LDS.128 along b, stores of 64-byte segments;
//   * mode 2 (contracted axis contiguous): lane <-> (b, output pair i = 2p, 2p+1), its 16 entries of M2 live in
//     registers; per plane a it reads row (a, b) (broadcast among the 4 lanes of the row) and writes 16 bytes of
//     it: every store instruction of the warp covers one whole 512-byte plane row block -- instead of the
//     row-per-thread form of fnsm_b200_tensor_product's mode 2, whose lanes write 64 bytes apart;
//   * results leave straight from registers with streaming stores.
#include "common.cuh"

namespace fnsm {

__device__ __forceinline__ void hd_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;"
               :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void hd_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void hd_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// streaming 16-byte store WITHOUT a "memory" clobber: the outputs are write-only, so the compiler may move the
// shared-memory loads of the next output row above it (with the clobber every store was a barrier and each row of
// the operator waited out its LDS latency: 8.5 TFLOP/s of DFMA, short-scoreboard 47 % -- profiles/r02_ncu_hexd_p7.txt)
__device__ __forceinline__ void hd_store(double* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v.x), "d"(v.y));
}

constexpr int kHdPlane = 72;                 // doubles between the a-planes of a staged element
constexpr int kHdElem = 8 * kHdPlane;        // 576 doubles = 4608 B per buffer
constexpr int kHdWarps = 8;

template <int NBUF>
__global__ void __launch_bounds__(kHdWarps * 32, 2)   // <= 128 registers: two CTAs (16 warps) per SM -- at 144 registers one CTA fit and the kernel sat at 71 % of DRAM peak
k_hex_deriv8(const double* __restrict__ A, const double* __restrict__ M0, const double* __restrict__ M1,
             const double* __restrict__ M2, double* __restrict__ o0, double* __restrict__ o1,
             double* __restrict__ o2, long long E, int skip) {   // skip: profiling aid, bit k = leave mode k out
  extern __shared__ __align__(16) double hd_smem[];
  double* sM = hd_smem;                                        // [2][64]: M0, M1 row-major [i][a]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = hd_smem + 128 + (size_t)warp * NBUF * kHdElem;
  for (int k = threadIdx.x; k < 128; k += blockDim.x) sM[k] = k < 64 ? M0[k] : M1[k - 64];
  // mode 2: this lane's two rows of M2
  const int p2 = lane & 3, rg = lane >> 2;
  double m2[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int c = 0; c < 8; ++c) m2[h][c] = M2[(2 * p2 + h) * 8 + c];
  __syncthreads();

  const long long wstride = (long long)gridDim.x * kHdWarps;
  const long long e_first = (long long)blockIdx.x * kHdWarps + warp;
  auto fetch = [&](long long e, int buf) {
    if (e < E) {
      const double* src = A + e * 512 + 2 * lane;
      double* dst = ring + buf * kHdElem + 2 * lane;
#pragma unroll
      for (int a = 0; a < 8; ++a) hd_cp_async16(dst + a * kHdPlane, src + a * 64);
    }
    hd_commit();                                               // (possibly empty) group: keeps the group count uniform
  };
#pragma unroll
  for (int k = 0; k < NBUF - 1; ++k) fetch(e_first + k * wstride, k);

  int buf = 0;
  for (long long e = e_first; e < E; e += wstride) {
    int nb = buf + NBUF - 1;
    if (nb >= NBUF) nb -= NBUF;
    fetch(e + (NBUF - 1) * wstride, nb);                       // its previous occupant was consumed last iteration
    hd_wait<NBUF - 1>();
    __syncwarp();
    const double* s = ring + buf * kHdElem;
    // ---- mode 0: d0[i][bc] = sum_a M0[i][a] A[a][bc], lane <-> bc = 2 lane, 2 lane + 1 ----
    if (!(skip & 1)) {
      double2 v[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) v[a] = *reinterpret_cast<const double2*>(s + a * kHdPlane + 2 * lane);
      double* o = o0 + e * 512 + 2 * lane;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int a = 0; a < 8; a += 2) {
          const double2 m = *reinterpret_cast<const double2*>(sM + i * 8 + a);
          acc.x = fma(m.x, v[a].x, acc.x); acc.y = fma(m.x, v[a].y, acc.y);
          acc.x = fma(m.y, v[a + 1].x, acc.x); acc.y = fma(m.y, v[a + 1].y, acc.y);
        }
        hd_store(o + i * 64, acc);
      }
    }
    // ---- mode 1: d1[a][i][c] = sum_b M1[i][b] A[a][b][c], lane <-> (a = lane / 4, c = 2 (lane % 4) ..+1) ----
    if (!(skip & 2)) {
      const int a = lane >> 2, cp = lane & 3;
      double2 v[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) v[b] = *reinterpret_cast<const double2*>(s + a * kHdPlane + b * 8 + 2 * cp);
      double* o = o1 + e * 512 + a * 64 + 2 * cp;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
          const double2 m = *reinterpret_cast<const double2*>(sM + 64 + i * 8 + b);
          acc.x = fma(m.x, v[b].x, acc.x); acc.y = fma(m.x, v[b].y, acc.y);
          acc.x = fma(m.y, v[b + 1].x, acc.x); acc.y = fma(m.y, v[b + 1].y, acc.y);
        }
        hd_store(o + i * 8, acc);
      }
    }
    // ---- mode 2: d2[a][b][i] = sum_c M2[i][c] A[a][b][c], lane <-> (b = lane / 4, i = 2 (lane % 4) ..+1) ----
    if (!(skip & 4)) {
      double* o = o2 + e * 512 + rg * 8 + 2 * p2;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const double* row = s + a * kHdPlane + rg * 8;
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          const double2 x = *reinterpret_cast<const double2*>(row + c);
          acc.x = fma(m2[0][c], x.x, acc.x); acc.y = fma(m2[1][c], x.x, acc.y);
          acc.x = fma(m2[0][c + 1], x.y, acc.x); acc.y = fma(m2[1][c + 1], x.y, acc.y);
        }
        hd_store(o + a * 64, acc);
      }
    }
    __syncwarp();                                              // every lane is done reading this buffer
    if (++buf == NBUF) buf = 0;
  }
  hd_wait<0>();
}

template <int NBUF>
static int launch_hex_deriv(const double* A, const double* const* M, double* const* outs, long long E,
                            const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  const size_t smem = 8 * (128 + (size_t)kHdWarps * NBUF * kHdElem);
  auto kernel = k_hex_deriv8<NBUF>;
  static std::atomic<int> occ_cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int occ = occ_cache[dev & 63].load(std::memory_order_relaxed);
  if (occ == 0) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return FNSM_E_BAD_CONFIG;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kHdWarps * 32, smem) != cudaSuccess || occ < 1) occ = 1;
    occ_cache[dev & 63].store(occ, std::memory_order_relaxed);
  }
  if (cfg && cfg->ctas_per_sm > 0 && cfg->ctas_per_sm < occ) occ = cfg->ctas_per_sm;
  long long grid = (long long)occ * di.sms;
  const long long need = (E + kHdWarps - 1) / kHdWarps;
  if (grid > need) grid = need;
  kernel<<<(unsigned)grid, kHdWarps * 32, smem, st>>>(A, M[0], M[1], M[2], outs[0], outs[1], outs[2], E,
                                                        cfg ? (cfg->reserved[0] >> 4) & 7 : 0);
  return post_launch();
}

}  // namespace fnsm

extern "C" int fnsm_b200_hex_deriv(int32_t dtype, const void* A, const void* const* M, void* const* outs,
                                   int32_t n1d, int64_t E, const fnsm_cfg* cfg, void* stream) {
  using namespace fnsm;
  if (!A || !M || !outs || E < 0) return FNSM_E_BAD_ARG;
  for (int k = 0; k < 3; ++k)
    if (!M[k] || !outs[k]) return FNSM_E_BAD_ARG;
  if (dtype != FNSM_F64 || n1d != 8) return FNSM_E_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if ((reinterpret_cast<uintptr_t>(outs[k]) & 15) != 0) return FNSM_E_ALIGNMENT;
  if ((reinterpret_cast<uintptr_t>(A) & 15) != 0) return FNSM_E_ALIGNMENT;
  if (E == 0) return FNSM_OK;
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double* const* Md = reinterpret_cast<const double* const*>(M);
  double* const* od = reinterpret_cast<double* const*>(outs);
  const int stages = (cfg && cfg->stages > 0) ? cfg->stages : 3;
  switch (stages) {
    case 2: return launch_hex_deriv<2>(static_cast<const double*>(A), Md, od, E, cfg, di, st);
    case 3: return launch_hex_deriv<3>(static_cast<const double*>(A), Md, od, E, cfg, di, st);
    case 4: return launch_hex_deriv<4>(static_cast<const double*>(A), Md, od, E, cfg, di, st);
    default: return FNSM_E_BAD_CONFIG;
  }
}
