// The shared-operator family   se,sij,ej->ei :   out_b[e,i] = sum_{s,j} R[s,i,j] J_b[s,e] u_b[e,j]
// (reference test/test_codegen.py:34-88 "div components" J{x,y,z}(3,E) R(3,35,35) u{x,y,z}(E,35) and "face mass"
// J(4,E) R(4,15,15) v_k(E,15); tuning/impls/re_rij_ej_to_ei*.py, re_rji_ej_to_ei_3d_cross_product_v0.py:78-89).
// Every row of the batch has its own geometric factors J_b and field u_b; the operator R is shared.
//
// Flop-optimal order = the face-mass lift's: scale first, w[s,e,j] = J[s,e] u[e,j] (S J multiplies), then ONE dense
// contraction over k = (s, j) against the resident operator (2 S I J flops) -- the reference's transforms instead
// compute T[s] = R[s] u and reuse it ("D u is reused"), which needs the J-weighted sum afterwards; both are one read
// of u per row.  On the FP64 tensor path that is the divergence kernel without its x-sum:
//   * tets p = 4 volume operator (S = 3, I = J = 35): k_div_dmma<.., NX = 1> (opmat_dmma.cuh) -- TMA producer,
//     27 k-tiles, dofs 32..34 on DFMA, direct stores;
//   * other shapes (S <= 4): the warp-per-chunk kernel below, in the style of opmat_dmma_gen.cuh (software-pipelined
//     plain loads, any E and alignment, operator table in fragment order, staged coalesced stores).
#pragma once
#include "opmat_dmma_gen.cuh"

namespace fnsm {

struct SeRows {
  const void* jac[8];
  const void* field[8];
  void* out[8];
};

template <int NS, int NI, int NJ>
struct SeLayout {
  static constexpr int ME = NJ <= 6 ? 4 : 2, CH = 8 * ME;
  static constexpr int JQ = (NJ + 3) / 4;
  static constexpr int KT = NS * JQ;                         // k-tile = (jq, s), k-in-tile t <-> j = 4 jq + t
  static constexpr int NT = (NI + 7) / 8;
  static constexpr int PITCH = 8 * NT + 2;
  static constexpr int B_DOUBLES = KT * NT * 32;
  static constexpr int IN_DOUBLES = CH * NJ, J_DOUBLES = NS * CH;
  static constexpr int SLOT_DOUBLES = IN_DOUBLES + J_DOUBLES;
  static constexpr int WARP_DOUBLES = SLOT_DOUBLES + CH * PITCH;
  static constexpr int NW = 8;
  static constexpr size_t SMEM = 8 * ((size_t)B_DOUBLES + (size_t)NW * WARP_DOUBLES);
};

// ES: geometric factors laid out J(E, S) ("es,sij,ej->ei") instead of J(S, E)
template <int NS, int NI, int NJ, bool ES>
__global__ void __launch_bounds__(256, 2)
k_se_dmma_gen(const double* __restrict__ Og, const __grid_constant__ SeRows rows, int nrows, long long E) {
  using L = SeLayout<NS, NI, NJ>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double* s = sB + L::B_DOUBLES + (size_t)warp * L::WARP_DOUBLES;
  double* sJ = s + L::IN_DOUBLES;
  double* stage = s + L::SLOT_DOUBLES;

  // operator table in fragment order: sB[(kt*NT + nt)*32 + lane] = R[s][n = 8 nt + g][j = 4 jq + t], kt = NS jq + s
  _Pragma("unroll 4")
  for (int idx = threadIdx.x; idx < L::B_DOUBLES; idx += blockDim.x) {
    const int ln = idx & 31, nt = (idx >> 5) % L::NT, kt = (idx >> 5) / L::NT;
    const int gg = ln >> 2, tt = ln & 3, n = 8 * nt + gg;
    const int jq = kt / NS, sx = kt - NS * jq, j = 4 * jq + tt;
    sB[idx] = (j < NJ && n < NI) ? Og[(sx * NI + n) * NJ + j] : 0.0;
  }
  __syncthreads();

  const long long nchunks = (E + L::CH - 1) / L::CH;
  const long long wstride = (long long)gridDim.x * L::NW;
  const long long chunk0 = (long long)blockIdx.x * L::NW + warp;
  const long long my_chunks = chunk0 < nchunks ? (nchunks - chunk0 + wstride - 1) / wstride : 0;
  const long long nitems = my_chunks * nrows;             // item = (chunk, row of the batched einsum)

  // software pipeline: item n + 1 travels global -> registers while item n is computed
  constexpr int NIN = (L::IN_DOUBLES + 31) / 32, NJR = (L::J_DOUBLES + 31) / 32;
  double rin[NIN], rj[NJR];
  // (chunk, row) of an item advance by counting: a 64-bit division per item costs as much as its DMMAs
  auto fetch = [&](long long chunk, int row) {
    const long long e0 = chunk * L::CH;
    const int ne = (int)((E - e0 < L::CH) ? (E - e0) : L::CH);
    const double* __restrict__ in = static_cast<const double*>(rows.field[row]);
    const double* __restrict__ Jg = static_cast<const double*>(rows.jac[row]);
#pragma unroll
    for (int q = 0; q < NIN; ++q) {
      const int k = lane + 32 * q;
      rin[q] = k < ne * NJ ? ldg_stream(in + e0 * NJ + k) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < NJR; ++q) {
      const int k = lane + 32 * q;
      if (ES) {
        rj[q] = k < ne * NS ? ldg_stream(Jg + e0 * NS + k) : 0.0;                 // contiguous [el][s]
      } else {
        const int sx = k / L::CH, el = k - sx * L::CH;
        rj[q] = (k < L::J_DOUBLES && el < ne) ? ldg_stream(Jg + (long long)sx * E + e0 + el) : 0.0;
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
#pragma unroll
    for (int q = 0; q < NIN; ++q)
      if (lane + 32 * q < L::IN_DOUBLES) s[lane + 32 * q] = rin[q];
#pragma unroll
    for (int q = 0; q < NJR; ++q)
      if (lane + 32 * q < L::J_DOUBLES) sJ[lane + 32 * q] = rj[q];
    if (item + 1 < nitems) fetch(chunk_n, row_n);
    __syncwarp();
    // ---- A fragments: lane (g, t) holds rows el = g + 8 m, k = (kt, t): w = J[s][el] u[el][j] ----
    double a[L::ME][L::KT];
#pragma unroll
    for (int m = 0; m < L::ME; ++m) {
      const int el = g + 8 * m;
      double Jr[NS];
#pragma unroll
      for (int sx = 0; sx < NS; ++sx) Jr[sx] = ES ? sJ[el * NS + sx] : sJ[sx * L::CH + el];
#pragma unroll
      for (int jq = 0; jq < L::JQ; ++jq) {
        const int j = 4 * jq + t;
        const double u = j < NJ ? s[el * NJ + j] : 0.0;
#pragma unroll
        for (int sx = 0; sx < NS; ++sx) a[m][NS * jq + sx] = Jr[sx] * u;
      }
    }
    // ---- DMMAs, accumulators -> stage[el][n] ----
#pragma unroll
    for (int nt0 = 0; nt0 < L::NT; nt0 += 4) {
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
    for (int idx = lane; idx < ne * NI; idx += 32) {
      const int el = idx / NI, i = idx - el * NI;
      stg_stream(out + e0 * NI + idx, stage[el * L::PITCH + i]);
    }
    __syncwarp();                                         // slot and stage are rewritten by the next item
    chunk = chunk_n;
    row = row_n;
  }
}

// shapes with a tensor-path instantiation (S, I, J): tets p = 1..4 volume operators, face operators, p = 4 faces
inline bool se_supported(int dtype, int ns, int ni, int nj) {
  if (dtype != FNSM_F64) return false;
  if (ns == 3 && ni == nj) return ni == 4 || ni == 10 || ni == 20 || ni == 35;
  if (ns == 4 && ni == nj) return ni == 3 || ni == 6 || ni == 10 || ni == 15;
  return false;
}

template <int NS, int NI, int NJ, bool ES>
static int launch_se_gen_es(const void* op, const SeRows& rows, int nrows, long long E, const fnsm_cfg* cfg,
                            const DevInfo& di, cudaStream_t st) {
  using L = SeLayout<NS, NI, NJ>;
  auto kernel = k_se_dmma_gen<NS, NI, NJ, ES>;
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
  kernel<<<(unsigned)grid, 256, L::SMEM, st>>>(static_cast<const double*>(op), rows, nrows, E);
  return post_launch();
}

template <int NS, int NI, int NJ>
static int launch_se_gen_k(const void* op, const SeRows& rows, int nrows, long long E, const fnsm_cfg* cfg,
                           const DevInfo& di, cudaStream_t st, bool es) {
  return es ? launch_se_gen_es<NS, NI, NJ, true>(op, rows, nrows, E, cfg, di, st)
            : launch_se_gen_es<NS, NI, NJ, false>(op, rows, nrows, E, cfg, di, st);
}

// p = 4 volume operator: the tuned divergence kernel with NX = 1, one launch per row (rows differ in J and u)
template <bool ES>
static int launch_se_p4(const void* op, const SeRows& rows, int nrows, long long E, const fnsm_cfg* cfg,
                        const DevInfo& di, cudaStream_t st) {
  constexpr int NW = 12;
  using L = DivLayoutT<1>;
  // with a third of the divergence's slot there is room for the output stage (one TMA store per chunk instead of
  // 18 streaming stores per lane): ptxas then fits the loop into 168 registers with 8 bytes of spills instead of 124
  constexpr bool STAGED = true;
  // (sized for the TMA = false instantiation, whose slots carry two doubles of headroom per slab)
  const size_t smem = 8 * ((size_t)L::B_DOUBLES + (size_t)NW * (L::SLOT_DOUBLES_BULK + (STAGED ? OUT_BLOCK : 0))) + 8 * (size_t)NW + 8;
  if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  auto kernel = k_div_dmma<NW, STAGED, 0, 1, ES, true>;
  auto kernel_plain = k_div_dmma<NW, STAGED, 0, 1, ES, false>;
  if (int rc = set_smem(kernel, smem)) return rc;
  if (int rc = set_smem(kernel_plain, smem)) return rc;
  const long long nchunks = (E + kCH - 1) / kCH;
  const long long need = (nchunks + NW - 1) / NW;
  const unsigned grid = (unsigned)(di.sms < need ? di.sms : need);
  const bool force_plain = cfg && (cfg->reserved[0] & 1);
  for (int r = 0; r < nrows; ++r) {
    const double* J = static_cast<const double*>(rows.jac[r]);
    const double* u = static_cast<const double*>(rows.field[r]);
    double* out = static_cast<double*>(rows.out[r]);
    OpMaps maps{};
    const bool tma = !force_plain && E % 2 == 0 && E < (1LL << 31) - kCH && aligned16(J) && aligned16(u) && aligned16(out) &&
                     (ES ? map_rows(&maps.jac, J, E, 3) : map_erows(&maps.jac, J, E, 3)) &&
                     map_rows(&maps.in, u, E, 35) && map_rows(&maps.out, out, E, 35);
    launch_k(tma ? kernel : kernel_plain, grid, NW * 32, smem, st, maps, J, static_cast<const double*>(op), u, out,
             (long long)E, tma ? kFlagTma : 0);
    if (int rc = post_launch()) return rc;
  }
  return FNSM_OK;
}

static int launch_se(int dtype, int es, const void* const* jacs, const void* op, const void* const* fields,
                     void* const* outs, int b, int ns, int ni, int nj, long long E, const fnsm_cfg* cfg,
                     const DevInfo& di, cudaStream_t st) {
  if (!se_supported(dtype, ns, ni, nj)) return FNSM_E_UNSUPPORTED;
  if (cfg && cfg->variant > 1) return FNSM_E_UNSUPPORTED;      // 0 = auto, 1 = FP64 tensor path; there is no simt / tcgen05 form
  for (int r0 = 0; r0 < b; r0 += 8) {
    const int nr = (b - r0 < 8) ? (b - r0) : 8;
    SeRows rows{};
    for (int r = 0; r < nr; ++r) {
      rows.jac[r] = jacs[r0 + r]; rows.field[r] = fields[r0 + r]; rows.out[r] = outs[r0 + r];
      if (!rows.jac[r] || !rows.field[r] || !rows.out[r]) return FNSM_E_BAD_ARG;
    }
    int rc = FNSM_E_UNSUPPORTED;
    if (ns == 3) {
      switch (ni) {
        case 4: rc = launch_se_gen_k<3, 4, 4>(op, rows, nr, E, cfg, di, st, es != 0); break;
        case 10: rc = launch_se_gen_k<3, 10, 10>(op, rows, nr, E, cfg, di, st, es != 0); break;
        case 20: rc = launch_se_gen_k<3, 20, 20>(op, rows, nr, E, cfg, di, st, es != 0); break;
        case 35: rc = es ? launch_se_p4<true>(op, rows, nr, E, cfg, di, st) : launch_se_p4<false>(op, rows, nr, E, cfg, di, st); break;
      }
    } else {
      switch (ni) {
        case 3: rc = launch_se_gen_k<4, 3, 3>(op, rows, nr, E, cfg, di, st, es != 0); break;
        case 6: rc = launch_se_gen_k<4, 6, 6>(op, rows, nr, E, cfg, di, st, es != 0); break;
        case 10: rc = launch_se_gen_k<4, 10, 10>(op, rows, nr, E, cfg, di, st, es != 0); break;
        case 15: rc = launch_se_gen_k<4, 15, 15>(op, rows, nr, E, cfg, di, st, es != 0); break;
      }
    }
    if (rc) return rc;
  }
  return FNSM_OK;
}

}  // namespace fnsm
