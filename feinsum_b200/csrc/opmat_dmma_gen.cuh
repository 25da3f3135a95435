// fp64 DMMA kernels for the lower-order tets (p = 1..3: 4 / 10 / 20 volume dofs, 3 / 6 / 10 face dofs).
//
// These shapes are HBM bound (arithmetic intensity 0.8 .. 3.9 flop/B against a ridge of 5.7), but p = 2
// and p = 3 still need 13 / 25 TFLOP/s of FP64 to keep up with HBM -- more than DFMA fed from shared
// memory delivers -- so the contraction runs on the FP64 tensor path (mma.sync.m8n8k4.f64) exactly like
// the p = 4 kernels of opmat_dmma.cuh, without their shape-specific layout tricks:
//   * a warp owns a chunk of 16 elements: coalesced loads into its shared-memory slot, A fragments
//     (Jacobian folded in / face Jacobian applied) built in registers, DMMAs against the operator
//     table held in shared memory in fragment order, accumulators staged in shared memory,
//   * grad applies J to the staged (dof, r) triples in a second pass, then every kernel leaves with
//     coalesced stores,
//   * plain loads and stores (any E, any alignment); the many small independent warps (8 per CTA,
//     several CTAs per SM) keep enough bytes in flight.
#pragma once
#include "opmat_dmma.cuh"

namespace fnsm {

constexpr int gen_pad(int v, int m) { return (v + m - 1) / m * m; }

// KIND: FNSM_OP_GRAD / FNSM_OP_DIV / FNSM_OP_LIFT_EF / FNSM_OP_LIFT_FE
template <int KIND, int ND, int NFD>
struct GenLayout {
  static constexpr bool GRAD = KIND == FNSM_OP_GRAD, DIV = KIND == FNSM_OP_DIV, LIFT = !GRAD && !DIV;
  // element tiles (of 8) per chunk: the small orders take 32 elements per item to amortise the per-item instructions
  static constexpr int ME = ND <= 4 ? 4 : 2, CH = 8 * ME;   // (p = 2 with 32 elements: 254 registers, grad 82 -> 77 %)
  static constexpr int JQ = (ND + 3) / 4;                                   // j-quads of a dof row
  // contraction: grad k = j; div k-tile = (jq, r), k-in-tile t <-> j = 4 jq + t; lift k = NFD f + j
  static constexpr int KT = GRAD ? JQ : (DIV ? 3 * JQ : gen_pad(4 * NFD, 4) / 4);
  static constexpr int N = GRAD ? 3 * ND : ND;                              // grad: column n = 3 i + r
  static constexpr int NT = (N + 7) / 8;
  static constexpr int PITCH = 8 * NT + 2;                                  // stage row pitch (doubles), skewed
  static constexpr int B_DOUBLES = KT * NT * 32;
  // slot: element data of one chunk, then the Jacobian entries [xr][el] (grad / div: 9, lift: 4)
  static constexpr int IN_DOUBLES = GRAD ? CH * ND : (DIV ? 3 * CH * ND : 4 * CH * NFD);
  static constexpr int J_DOUBLES = (LIFT ? 4 : 9) * CH;
  static constexpr int SLOT_DOUBLES = IN_DOUBLES + J_DOUBLES;
  static constexpr int STAGE_DOUBLES = CH * PITCH;
  static constexpr int WARP_DOUBLES = SLOT_DOUBLES + STAGE_DOUBLES;
  static constexpr int NW = 8;
  static constexpr size_t SMEM = 8 * ((size_t)B_DOUBLES + (size_t)NW * WARP_DOUBLES);
};

template <int KIND, int ND, int NFD>
// two CTAs per SM (<= 128 registers) wherever that does not spill: div at p = 3 holds 60 A-fragment and 60 prefetch
// registers and runs as one CTA per SM
__global__ void __launch_bounds__(256, (KIND == FNSM_OP_DIV && ND >= 20) ? 1 : 2)
k_opmat_dmma_gen(const double* __restrict__ Jg, const double* __restrict__ Og, const __grid_constant__ OpmatRows rows,
                 int nrows, long long E) {
  using L = GenLayout<KIND, ND, NFD>;
  constexpr bool FE = KIND == FNSM_OP_LIFT_FE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double* s = sB + L::B_DOUBLES + (size_t)warp * L::WARP_DOUBLES;
  double* sJ = s + L::IN_DOUBLES;
  double* stage = s + L::SLOT_DOUBLES;

  // operator table in fragment order: sB[(kt*NT + nt)*32 + lane] = B[k = (kt, t)][n = 8 nt + g]
  _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
  for (int idx = threadIdx.x; idx < L::B_DOUBLES; idx += blockDim.x) {
    const int ln = idx & 31, nt = (idx >> 5) % L::NT, kt = (idx >> 5) / L::NT;
    const int gg = ln >> 2, tt = ln & 3, n = 8 * nt + gg;
    double v = 0.0;
    if (L::GRAD) {
      const int j = 4 * kt + tt, i = n / 3, r = n - 3 * i;
      if (j < ND && n < L::N) v = Og[(r * ND + i) * ND + j];
    } else if (L::DIV) {
      const int jq = kt / 3, r = kt - 3 * jq, j = 4 * jq + tt;
      if (j < ND && n < ND) v = Og[(r * ND + n) * ND + j];
    } else {
      const int k = 4 * kt + tt, f = k / NFD, j = k - NFD * f;
      if (k < 4 * NFD && n < ND) v = FE ? Og[(n * 4 + f) * NFD + j] : Og[(f * ND + n) * NFD + j];
    }
    sB[idx] = v;
  }
  __syncthreads();

  const long long nchunks = (E + L::CH - 1) / L::CH;
  const long long wstride = (long long)gridDim.x * L::NW;
  const long long chunk0 = (long long)blockIdx.x * L::NW + warp;
  const long long my_chunks = chunk0 < nchunks ? (nchunks - chunk0 + wstride - 1) / wstride : 0;
  const long long nitems = my_chunks * nrows;             // item = (chunk, row of the batched einsum)

  // Software pipeline: the element data (and, for the first row of a chunk, the Jacobian entries) of
  // item n + 1 are fetched into registers with coalesced loads while item n is being computed.
  constexpr int NIN = L::IN_DOUBLES / 32, NJR = (L::J_DOUBLES + 31) / 32;
  static_assert(L::IN_DOUBLES % 32 == 0, "slot size must be a multiple of the warp size");
  double rin[NIN], rj[NJR];
  // (chunk, row) of an item advance by counting: a 64-bit division per item costs as much as its DMMAs
  auto fetch = [&](long long chunk, int row) {
    const long long e0 = chunk * L::CH;
    const int ne = (int)((E - e0 < L::CH) ? (E - e0) : L::CH);
    const double* __restrict__ in = static_cast<const double*>(rows.field[row]);
    constexpr int W = L::GRAD || L::DIV ? ND : NFD;        // row length of one slab
#pragma unroll
    for (int q = 0; q < NIN; ++q) {
      const int k = lane + 32 * q, slab = k / (L::CH * W), kk = k - slab * (L::CH * W);
      rin[q] = kk < ne * W ? ldg_stream(in + ((long long)slab * E + e0) * W + kk) : 0.0;
    }
    if (row == 0) {
#pragma unroll
      for (int q = 0; q < NJR; ++q) {
        const int k = lane + 32 * q;
        double v = 0.0;
        if (L::LIFT && !FE) {
          if (k < L::J_DOUBLES && (k >> 2) < ne) v = ldg_stream(Jg + e0 * 4 + k);       // J(E, 4): contiguous
        } else {
          const int xr = k / L::CH, el = k - xr * L::CH;
          if (k < L::J_DOUBLES && el < ne) v = ldg_stream(Jg + (long long)xr * E + e0 + el);
        }
        rj[q] = v;
      }
    }
  };
  if (nitems > 0) fetch(chunk0, 0);
  long long chunk = chunk0;
  int row = 0;
  for (long long item = 0; item < nitems; ++item) {
    // the item after this one
    const int row_n = row + 1 == nrows ? 0 : row + 1;
    const long long chunk_n = row_n == 0 ? chunk + wstride : chunk;
    const long long e0 = chunk * L::CH;
    const int ne = (int)((E - e0 < L::CH) ? (E - e0) : L::CH);
    double* __restrict__ out = static_cast<double*>(rows.out[row]);
    {
      // ---- registers -> slot (the previous item's stores have been issued; its slot reads are done) ----
#pragma unroll
      for (int q = 0; q < NIN; ++q) s[lane + 32 * q] = rin[q];
      if (row == 0) {
#pragma unroll
        for (int q = 0; q < NJR; ++q) {
          const int k = lane + 32 * q;
          if (k < L::J_DOUBLES) {
            if (L::LIFT && !FE) sJ[(k & 3) * L::CH + (k >> 2)] = rj[q];                    // transposed to [f][el]
            else sJ[k] = rj[q];
          }
        }
      }
      if (item + 1 < nitems) fetch(chunk_n, row_n);
      __syncwarp();
      // ---- A fragments: lane (g, t) holds rows el = g + 8 m, k = (kt, t) ----
      double a[L::ME][L::KT];
#pragma unroll
      for (int m = 0; m < L::ME; ++m) {
        const int el = g + 8 * m;
        if (L::GRAD) {
#pragma unroll
          for (int kt = 0; kt < L::KT; ++kt) {
            const int j = 4 * kt + t;
            a[m][kt] = j < ND ? s[el * ND + j] : 0.0;
          }
        } else if (L::DIV) {
          double Jr[9];
#pragma unroll
          for (int xr = 0; xr < 9; ++xr) Jr[xr] = sJ[xr * L::CH + el];
#pragma unroll
          for (int jq = 0; jq < L::JQ; ++jq) {
            const int j = 4 * jq + t;
            double ux[3];
#pragma unroll
            for (int x = 0; x < 3; ++x) ux[x] = j < ND ? s[(x * L::CH + el) * ND + j] : 0.0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
              a[m][3 * jq + r] = fma(Jr[6 + r], ux[2], fma(Jr[3 + r], ux[1], Jr[r] * ux[0]));
          }
        } else {
#pragma unroll
          for (int kt = 0; kt < L::KT; ++kt) {
            const int k = 4 * kt + t, f = k / NFD, j = k - NFD * f;
            a[m][kt] = k < 4 * NFD ? sJ[f * L::CH + el] * s[(f * L::CH + el) * NFD + j] : 0.0;
          }
        }
      }
      // ---- DMMAs, accumulators -> stage[el][n] ----
#pragma unroll
      for (int nt0 = 0; nt0 < L::NT; nt0 += 4) {          // at most 4 column tiles (16 accumulators) at a time
        constexpr int NTG = 4;
        double acc[L::ME][NTG][2];
#pragma unroll
        for (int m = 0; m < L::ME; ++m)
#pragma unroll
          for (int q = 0; q < NTG; ++q) { acc[m][q][0] = 0.0; acc[m][q][1] = 0.0; }
#pragma unroll
        for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
          for (int q = 0; q < NTG; ++q) {
            if (nt0 + q < L::NT) {
              const double b = sB[(kt * L::NT + nt0 + q) * 32 + lane];
#pragma unroll
              for (int m = 0; m < L::ME; ++m) dmma884(acc[m][q], a[m][kt], b);
            }
          }
        }
#pragma unroll
        for (int m = 0; m < L::ME; ++m)
#pragma unroll
          for (int q = 0; q < NTG; ++q)
            if (nt0 + q < L::NT)
              *reinterpret_cast<double2*>(stage + (g + 8 * m) * L::PITCH + 8 * (nt0 + q) + 2 * t) =
                  make_double2(acc[m][q][0], acc[m][q][1]);
      }
      __syncwarp();
      // ---- coalesced stores (grad: J applied to the staged (dof, r) triples) ----
      for (int idx = lane; idx < ne * ND; idx += 32) {
        const int el = idx / ND, i = idx - el * ND;
        if (L::GRAD) {
          const double* T = stage + el * L::PITCH + 3 * i;
#pragma unroll
          for (int x = 0; x < 3; ++x)
            stg_stream(out + ((long long)x * E + e0) * ND + idx,
                       fma(sJ[(3 * x + 2) * L::CH + el], T[2], fma(sJ[(3 * x + 1) * L::CH + el], T[1], sJ[(3 * x) * L::CH + el] * T[0])));
        } else {
          stg_stream(out + e0 * ND + idx, stage[el * L::PITCH + i]);
        }
      }
      __syncwarp();                                       // slot and stage are rewritten by the next item
    }
    chunk = chunk_n;
    row = row_n;
  }
}

inline bool dmma_gen_supported(int kind, int n_outer, int ni, int nj) {
  if (kind == FNSM_OP_GRAD || kind == FNSM_OP_DIV) return n_outer == 3 && ni == nj && (ni == 4 || ni == 10 || ni == 20);
  return n_outer == 4 && ((ni == 4 && nj == 3) || (ni == 10 && nj == 6) || (ni == 20 && nj == 10));
}

template <int KIND, int ND, int NFD>
static int launch_dmma_gen_k(const void* jac, const void* op, const OpmatRows& rows, int nrows, long long E,
                             const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  using L = GenLayout<KIND, ND, NFD>;
  auto kernel = k_opmat_dmma_gen<KIND, ND, NFD>;
  if (L::SMEM > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  if (int rc = set_smem(kernel, L::SMEM)) return rc;
  static std::atomic<int> occ_cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int occ = occ_cache[dev & 63].load(std::memory_order_relaxed);
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, L::SMEM) != cudaSuccess || occ < 1) occ = 1;
    occ_cache[dev & 63].store(occ, std::memory_order_relaxed);
  }
  if (cfg && cfg->ctas_per_sm > 0 && cfg->ctas_per_sm < occ) occ = cfg->ctas_per_sm;
  const long long nchunks = (E + L::CH - 1) / L::CH;
  const long long need = (nchunks + L::NW - 1) / L::NW;
  long long grid = (long long)occ * di.sms;
  if (grid > need) grid = need;
  kernel<<<(unsigned)grid, 256, L::SMEM, st>>>(static_cast<const double*>(jac), static_cast<const double*>(op),
                                                 rows, nrows, E);
  return post_launch();
}

template <int ND, int NFD>
static int launch_dmma_gen_order(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                                 long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  switch (kind) {
    case FNSM_OP_GRAD: return launch_dmma_gen_k<FNSM_OP_GRAD, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    case FNSM_OP_DIV: return launch_dmma_gen_k<FNSM_OP_DIV, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    case FNSM_OP_LIFT_EF: return launch_dmma_gen_k<FNSM_OP_LIFT_EF, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    case FNSM_OP_LIFT_FE: return launch_dmma_gen_k<FNSM_OP_LIFT_FE, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    default: return FNSM_E_BAD_ARG;
  }
}

static int launch_dmma_gen(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows, int ni,
                           long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  switch (ni) {
    case 4: return launch_dmma_gen_order<4, 3>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 10: return launch_dmma_gen_order<10, 6>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 20: return launch_dmma_gen_order<20, 10>(kind, jac, op, rows, nrows, E, cfg, di, st);
    default: return FNSM_E_UNSUPPORTED;
  }
}

}  // namespace fnsm
