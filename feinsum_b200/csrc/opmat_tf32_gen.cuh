// fp32 companion of opmat_dmma_gen.cuh for the lower-order tets (p = 1..3): the same warp-per-chunk
// structure (coalesced plain loads with the next item prefetched into registers, A fragments in
// registers, operator table in shared memory in fragment order, staged accumulators, coalesced stores;
// any E, any alignment), with the contraction on the legacy tensor path as 3xTF32
// (mma.sync.m16n8k8.tf32, hi/lo split of both operands, cross terms in their own accumulator -- see
// opmat_tf32.cuh).  These shapes are HBM bound (at most 50 TFLOP/s fp32-equivalent are needed at p = 3,
// the 3xTF32 path delivers 92), so the kernel is judged by its HBM bandwidth.  A chunk of 16 elements
// is exactly one M tile.
#pragma once
#include "opmat_tf32.cuh"
#include "opmat_dmma_gen.cuh"

namespace fnsm {

template <int KIND, int ND, int NFD>
struct Gen32Layout {
  static constexpr bool GRAD = KIND == FNSM_OP_GRAD, DIV = KIND == FNSM_OP_DIV, LIFT = !GRAD && !DIV;
  static constexpr int JQ = (ND + 7) / 8;                                   // j-octets of a dof row
  // contraction: grad k = j; div k-tile = (jq, r), k-in-tile t (+4) <-> j = 8 jq + t (+4); lift k = NFD f + j
  static constexpr int KT = GRAD ? JQ : (DIV ? 3 * JQ : (4 * NFD + 7) / 8);
  static constexpr int N = GRAD ? 3 * ND : ND;                              // grad: column n = 3 i + r
  static constexpr int NT = (N + 7) / 8;
  static constexpr int PITCH = 8 * NT + 2;                                  // stage row pitch (floats), skewed
  static constexpr int B_BYTES = KT * NT * 32 * 16;                         // uint4 {hi, hi, lo, lo} per lane and fragment
  static constexpr int IN_FLOATS = GRAD ? kCH * ND : (DIV ? 3 * kCH * ND : 4 * kCH * NFD);
  static constexpr int J_FLOATS = (LIFT ? 4 : 9) * kCH;
  static constexpr int SLOT_FLOATS = IN_FLOATS + J_FLOATS;
  static constexpr int STAGE_FLOATS = kCH * PITCH;
  static constexpr int WARP_FLOATS = SLOT_FLOATS + STAGE_FLOATS;
  static constexpr int NW = 8;
  static constexpr size_t SMEM = (size_t)B_BYTES + 4 * (size_t)NW * WARP_FLOATS;
};

template <int KIND, int ND, int NFD>
__global__ void __launch_bounds__(256)
k_opmat_tf32_gen(const float* __restrict__ Jg, const float* __restrict__ Og, const __grid_constant__ OpmatRows rows,
                 int nrows, long long E) {
  using L = Gen32Layout<KIND, ND, NFD>;
  constexpr bool FE = KIND == FNSM_OP_LIFT_FE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* sB = reinterpret_cast<uint4*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* s = reinterpret_cast<float*>(smem_raw + L::B_BYTES) + (size_t)warp * L::WARP_FLOATS;
  float* sJ = s + L::IN_FLOATS;
  float* stage = s + L::SLOT_FLOATS;

  // operator table: fragment (kt, nt), lane (n = 8 nt + g, k = (kt, t) and (kt, t + 4)), split hi / lo
  fill_b_table(sB, L::KT * L::NT, [&](int frag, int ln, int half) {
    const int kt = frag / L::NT, nt = frag - kt * L::NT;
    const int gg = ln >> 2, tt = (ln & 3) + 4 * half, n = 8 * nt + gg;
    if (L::GRAD) {
      const int j = 8 * kt + tt, i = n / 3, r = n - 3 * i;
      return (j < ND && n < L::N) ? Og[(r * ND + i) * ND + j] : 0.f;
    } else if (L::DIV) {
      const int jq = kt / 3, r = kt - 3 * jq, j = 8 * jq + tt;
      return (j < ND && n < ND) ? Og[(r * ND + n) * ND + j] : 0.f;
    } else {
      const int k = 8 * kt + tt, f = k / NFD, j = k - NFD * f;
      if (k >= 4 * NFD || n >= ND) return 0.f;
      return FE ? Og[(n * 4 + f) * NFD + j] : Og[(f * ND + n) * NFD + j];
    }
  });
  __syncthreads();

  const long long nchunks = (E + kCH - 1) / kCH;
  const long long wstride = (long long)gridDim.x * L::NW;
  const long long chunk0 = (long long)blockIdx.x * L::NW + warp;
  const long long my_chunks = chunk0 < nchunks ? (nchunks - chunk0 + wstride - 1) / wstride : 0;
  const long long nitems = my_chunks * nrows;             // item = (chunk, row of the batched einsum)

  constexpr int NIN = L::IN_FLOATS / 32, NJR = (L::J_FLOATS + 31) / 32;
  static_assert(L::IN_FLOATS % 32 == 0, "slot size must be a multiple of the warp size");
  float rin[NIN], rj[NJR];
  // (chunk, row) of an item advance by counting: a 64-bit division per item costs as much as its DMMAs
  auto fetch = [&](long long chunk, int row) {
    const long long e0 = chunk * kCH;
    const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
    const float* __restrict__ in = static_cast<const float*>(rows.field[row]);
    constexpr int W = L::GRAD || L::DIV ? ND : NFD;        // row length of one slab
#pragma unroll
    for (int q = 0; q < NIN; ++q) {
      const int k = lane + 32 * q, slab = k / (kCH * W), kk = k - slab * (kCH * W);
      rin[q] = kk < ne * W ? ldg_stream(in + ((long long)slab * E + e0) * W + kk) : 0.f;
    }
    if (row == 0) {
#pragma unroll
      for (int q = 0; q < NJR; ++q) {
        const int k = lane + 32 * q;
        float v = 0.f;
        if (L::LIFT && !FE) {
          if (k < L::J_FLOATS && (k >> 2) < ne) v = ldg_stream(Jg + e0 * 4 + k);        // J(E, 4): contiguous
        } else {
          const int xr = k / kCH, el = k - xr * kCH;
          if (k < L::J_FLOATS && el < ne) v = ldg_stream(Jg + (long long)xr * E + e0 + el);
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
    const long long e0 = chunk * kCH;
    const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
    float* __restrict__ out = static_cast<float*>(rows.out[row]);
#pragma unroll
    for (int q = 0; q < NIN; ++q) s[lane + 32 * q] = rin[q];
    if (row == 0) {
#pragma unroll
      for (int q = 0; q < NJR; ++q) {
        const int k = lane + 32 * q;
        if (k < L::J_FLOATS) {
          if (L::LIFT && !FE) sJ[(k & 3) * kCH + (k >> 2)] = rj[q];                      // transposed to [f][el]
          else sJ[k] = rj[q];
        }
      }
    }
    if (item + 1 < nitems) fetch(chunk_n, row_n);
    __syncwarp();
    // ---- A fragments (m16n8k8): a[2c + h] = A[row g + 8h][k = (kt, t + 4c)], split hi / lo ----
    uint32_t ahi[L::KT][4], alo[L::KT][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int el = g + 8 * h;
      if (L::GRAD) {
#pragma unroll
        for (int kt = 0; kt < L::KT; ++kt)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int j = 8 * kt + t + 4 * c;
            split_tf32(j < ND ? s[el * ND + j] : 0.f, ahi[kt][2 * c + h], alo[kt][2 * c + h]);
          }
      } else if (L::DIV) {
        float Jr[9];
#pragma unroll
        for (int xr = 0; xr < 9; ++xr) Jr[xr] = sJ[xr * kCH + el];
#pragma unroll
        for (int jq = 0; jq < L::JQ; ++jq)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int j = 8 * jq + t + 4 * c;
            float ux[3];
#pragma unroll
            for (int x = 0; x < 3; ++x) ux[x] = j < ND ? s[(x * kCH + el) * ND + j] : 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r)
              split_tf32(fmaf(Jr[6 + r], ux[2], fmaf(Jr[3 + r], ux[1], Jr[r] * ux[0])),
                         ahi[3 * jq + r][2 * c + h], alo[3 * jq + r][2 * c + h]);
          }
      } else {
#pragma unroll
        for (int kt = 0; kt < L::KT; ++kt)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int k = 8 * kt + t + 4 * c, f = k / NFD, j = k - NFD * f;
            split_tf32(k < 4 * NFD ? sJ[f * kCH + el] * s[(f * kCH + el) * NFD + j] : 0.f,
                       ahi[kt][2 * c + h], alo[kt][2 * c + h]);
          }
      }
    }
    // ---- 3xTF32 MMAs, accumulators -> stage[el][n] ----
#pragma unroll
    for (int nt0 = 0; nt0 < L::NT; nt0 += 4) {            // at most 4 column tiles at a time
      constexpr int NTG = 4;
      float acc[NTG][4], corr[NTG][4];
#pragma unroll
      for (int q = 0; q < NTG; ++q)
#pragma unroll
        for (int v = 0; v < 4; ++v) { acc[q][v] = 0.f; corr[q][v] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt)
#pragma unroll
        for (int q = 0; q < NTG; ++q)
          if (nt0 + q < L::NT) mma_3xtf32(acc[q], corr[q], ahi[kt], alo[kt], sB[(kt * L::NT + nt0 + q) * 32 + lane]);
#pragma unroll
      for (int q = 0; q < NTG; ++q)
        if (nt0 + q < L::NT) {
          float* o = stage + g * L::PITCH + 8 * (nt0 + q) + 2 * t;
          *reinterpret_cast<float2*>(o) = make_float2(acc[q][0] + corr[q][0], acc[q][1] + corr[q][1]);
          *reinterpret_cast<float2*>(o + 8 * L::PITCH) = make_float2(acc[q][2] + corr[q][2], acc[q][3] + corr[q][3]);
        }
    }
    __syncwarp();
    // ---- coalesced stores (grad: J applied to the staged (dof, r) triples) ----
    for (int idx = lane; idx < ne * ND; idx += 32) {
      const int el = idx / ND, i = idx - el * ND;
      if (L::GRAD) {
        const float* T = stage + el * L::PITCH + 3 * i;
#pragma unroll
        for (int x = 0; x < 3; ++x)
          stg_stream(out + ((long long)x * E + e0) * ND + idx,
                     fmaf(sJ[(3 * x + 2) * kCH + el], T[2], fmaf(sJ[(3 * x + 1) * kCH + el], T[1], sJ[(3 * x) * kCH + el] * T[0])));
      } else {
        stg_stream(out + e0 * ND + idx, stage[el * L::PITCH + i]);
      }
    }
    __syncwarp();                                         // slot and stage are rewritten by the next item
    chunk = chunk_n;
    row = row_n;
  }
}

template <int KIND, int ND, int NFD>
static int launch_tf32_gen_k(const void* jac, const void* op, const OpmatRows& rows, int nrows, long long E,
                             const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  using L = Gen32Layout<KIND, ND, NFD>;
  auto kernel = k_opmat_tf32_gen<KIND, ND, NFD>;
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
  const long long nchunks = (E + kCH - 1) / kCH;
  const long long need = (nchunks + L::NW - 1) / L::NW;
  long long grid = (long long)occ * di.sms;
  if (grid > need) grid = need;
  kernel<<<(unsigned)grid, 256, L::SMEM, st>>>(static_cast<const float*>(jac), static_cast<const float*>(op),
                                                 rows, nrows, E);
  return post_launch();
}

template <int ND, int NFD>
static int launch_tf32_gen_order(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                                 long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  switch (kind) {
    case FNSM_OP_GRAD: return launch_tf32_gen_k<FNSM_OP_GRAD, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    case FNSM_OP_DIV: return launch_tf32_gen_k<FNSM_OP_DIV, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    case FNSM_OP_LIFT_EF: return launch_tf32_gen_k<FNSM_OP_LIFT_EF, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    case FNSM_OP_LIFT_FE: return launch_tf32_gen_k<FNSM_OP_LIFT_FE, ND, NFD>(jac, op, rows, nrows, E, cfg, di, st);
    default: return FNSM_E_BAD_ARG;
  }
}

static int launch_tf32_gen(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows, int ni,
                           long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  switch (ni) {
    case 4: return launch_tf32_gen_order<4, 3>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 10: return launch_tf32_gen_order<10, 6>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 20: return launch_tf32_gen_order<20, 10>(kind, jac, op, rows, nrows, E, cfg, di, st);
    default: return FNSM_E_UNSUPPORTED;
  }
}

}  // namespace fnsm
