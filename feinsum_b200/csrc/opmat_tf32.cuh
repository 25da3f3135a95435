// fp32 variant of the DG operator kernels (p = 4 tets): 3xTF32 on the legacy
// tensor path (mma.sync.m16n8k8.tf32 -> SASS HMMA.1688.F32.TF32).
//
// Plain TF32 (10-bit mantissa) breaks the 1e-5 tolerance of the fp32 einsums
// (SURVEY.md section 7), and the FFMA roofline of these kernels (7 980 flop per
// 596 B element) sits at the FP32 CUDA-core peak, which a kernel that must feed
// every FFMA from shared memory cannot approach.  Splitting both operands into
// hi = tf32(x), lo = tf32(x - hi) and accumulating a_lo*b_hi + a_hi*b_lo +
// a_hi*b_hi in fp32 restores ~2^-21 relative accuracy per product (the lo*lo term
// is dropped) at 3 tensor instructions per product: measured 277 TFLOP/s TF32
// on this path (tools/ubench4) -> 92 TFLOP/s fp32-equivalent, above the
// 73.5 TFLOP/s FFMA2 peak, so the kernels end up close to their HBM time.
//
// Same structure as opmat_dmma.cuh (warp-private slot, A fragments in registers,
// slot re-armed with the next item's TMA tensor loads, staged TMA store, CTA-local
// work queue).  M = 16 elements (one chunk = one M tile), N = 8 columns, K = 8.
// fp32 rows of 35 / 15 values are 140 / 60 B, so the tensor maps view the element
// axis in QUADS (560 / 240 B rows): the TMA path needs E % 4 == 0.
#pragma once
#include "opmat_dmma.cuh"

namespace fnsm {

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (c + corr) += (ahi + alo) * (bhi + blo) without the lo*lo term.  The tensor core adds into the fp32
// accumulator with truncation, a bias of up to one ulp per accumulate step that all-positive data
// does not average out; keeping the two small cross terms in their own accumulator cuts the steps
// on the large one from 3 to 1 per k-tile (measured bias vs numpy fp32: 1.9e-6 -> see tests).
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], float (&corr)[4], const uint32_t (&ahi)[4],
                                           const uint32_t (&alo)[4], uint2 bhi, uint2 blo) {
  mma_tf32(corr, alo, bhi.x, bhi.y);
  mma_tf32(corr, ahi, blo.x, blo.y);
  mma_tf32(c, ahi, bhi.x, bhi.y);
}

constexpr int OUT_BLOCK32 = kCH * 35;                // floats of one [16][35] output block (2240 B)
constexpr int align128(int bytes) { return (bytes + 127) / 128 * 128; }

__device__ __forceinline__ void flush_plain32(float* __restrict__ dst, const float* stage, long long e0,
                                              long long E, int lane) {
  const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
  for (int k = lane; k < ne * 35; k += 32) dst[k] = stage[k];
}

// fragment table: operator split into hi / lo, one uint4 {hi(k = t), hi(k = t + 4), lo(k = t), lo(k = t + 4)}
// per lane and fragment -> a single conflict-free LDS.128 feeds the three MMAs of a product
template <class F>
__device__ __forceinline__ void fill_b_table(uint4* tab, int n_frag, F value_at /* (frag, lane, half) */) {
  _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
  for (int idx = threadIdx.x; idx < n_frag * 32; idx += blockDim.x) {
    const int frag = idx >> 5, ln = idx & 31;
    uint4 v;
    split_tf32(value_at(frag, ln, 0), v.x, v.z);
    split_tf32(value_at(frag, ln, 1), v.y, v.w);
    tab[idx] = v;
  }
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], float (&corr)[4], const uint32_t (&ahi)[4],
                                           const uint32_t (&alo)[4], uint4 b) {
  mma_tf32(corr, alo, b.x, b.y);
  mma_tf32(corr, ahi, b.z, b.w);
  mma_tf32(c, ahi, b.x, b.y);
}

// ================================================================= DIV =====
// out[e,i] = sum_k B[k][i] w[e][k],  k = 35 r + j (105 -> 112),  w = sum_x J[x,r,e] u[x,e,j]
struct Div32 {
  static constexpr int KT = 14, NT = 5;
  static constexpr int B_BYTES = 2 * KT * NT * 32 * 8;                  // hi + lo tables
  static constexpr int U_SLAB = kCH * 35;                               // floats per x
  static constexpr int J_OFF = align128(3 * U_SLAB * 4) / 4;            // J region (floats), 128-B aligned for TMA
  static constexpr int SLOT_BYTES_TX = (3 * U_SLAB + 9 * kCH) * 4;      // bytes the two TMA loads deliver
  static constexpr int SLOT_BYTES = align128(J_OFF * 4 + 9 * kCH * 4);
  static constexpr int STAGE_BYTES = align128(OUT_BLOCK32 * 4);
};

__device__ __forceinline__ void div32_issue(float* s, uint64_t* bar, const OpMaps* maps, const float* __restrict__ Jg,
                                            const float* __restrict__ ug, long long chunk, long long E, bool tma, int lane) {
  using L = Div32;
  const long long e0 = chunk * kCH;
  if (tma) {
    if (elect_one()) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, L::SLOT_BYTES_TX);
      tma_load_3d(s, &maps->in, 0, (int)(chunk * (kCH / 4)), 0, bar);
      tma_load_2d(s + L::J_OFF, &maps->jac, (int)e0, 0, bar);
    }
  } else {
    const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
    for (int x = 0; x < 3; ++x)
      for (int k = lane; k < L::U_SLAB; k += 32)
        s[x * L::U_SLAB + k] = (k < ne * 35) ? ug[((long long)x * E + e0) * 35 + k] : 0.f;
    for (int k = lane; k < 9 * kCH; k += 32) {
      const int xr = k / kCH, el = k - xr * kCH;
      s[L::J_OFF + k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : 0.f;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_div_tf32(const __grid_constant__ OpMaps maps, const float* __restrict__ Jg, const float* __restrict__ Dg,
           const float* __restrict__ ug, float* __restrict__ outg, long long E, int flags) {
  using L = Div32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* sB = reinterpret_cast<uint4*>(smem_raw);
  unsigned char* slots = smem_raw + L::B_BYTES;
  unsigned char* stages = slots + (size_t)NW * L::SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)NW * L::STAGE_BYTES);
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], 1);
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();

  float* s = reinterpret_cast<float*>(slots + (size_t)warp * L::SLOT_BYTES);
  float* stage = reinterpret_cast<float*>(stages + (size_t)warp * L::STAGE_BYTES);
  uint64_t* bar = &bars[warp];
  const float* sJ = s + L::J_OFF;
  const long long nchunks = (E + kCH - 1) / kCH;
  const WorkQueue wq{work_ctr, nchunks};
  const int g = lane >> 2, t = lane & 3;
  const bool tma = flags & kFlagTma;

  long long cur = wq.take(lane), nxt = wq.take(lane);
  if (cur < nchunks) div32_issue(s, bar, &maps, Jg, ug, cur, E, tma, lane);
  // operator tables are staged while the first TMA loads are in flight
  // fragment (kt, nt): lane (n = g, k = t / t + 4): D[r][8nt+g][j], 35 r + j = 8kt + t (+4)
  fill_b_table(sB, L::KT * L::NT, [&](int frag, int ln, int half) {
    const int kt = frag / L::NT, nt = frag - kt * L::NT;
    const int g = ln >> 2, t = ln & 3, k = 8 * kt + t + 4 * half, i = 8 * nt + g;
    const int r = k / 35, j = k - 35 * r;
    return (k < 105 && i < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.f;
  });
  __syncthreads();
  for (uint32_t n = 0; cur < nchunks; ++n) {
    mbar_wait(bar, n & 1u);
    // ---- slot -> A fragments: rows g and g + 8 of the chunk, Jacobian folded in, split hi / lo ----
    uint32_t ahi[L::KT][4], alo[L::KT][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {                      // h = 0: element g, h = 1: element g + 8
      const int el = g + 8 * h;
      float Jr[9];
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jr[xr] = sJ[xr * kCH + el];
      const float* su = s + el * 35 + t;
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {                  // column t (c = 0) and t + 4 (c = 1) of the k-tile
          // k = 8kt + 4c + t = 35 r + j: r = R0 below the straddle threshold, R0 + 1 at / above it
          const int K0 = 8 * kt + 4 * c, R0 = K0 / 35, THR = 35 * (R0 + 1) - K0;
          float w = 0.f;
          if (K0 < 105) {
            const int off0 = K0 - 35 * R0;             // j for t = 0 and r = R0
            if (THR < 4) {
              const bool up = t >= THR;
              const int RU = R0 + 1 < 3 ? R0 + 1 : 2;
              const bool valid = !up || R0 + 1 < 3;    // k >= 105 is padding
              const int off = up ? off0 - 35 : off0;
              const float u0 = su[off], u1 = su[L::U_SLAB + off], u2 = su[2 * L::U_SLAB + off];
              const float j0 = up ? Jr[RU] : Jr[R0], j1 = up ? Jr[3 + RU] : Jr[3 + R0], j2 = up ? Jr[6 + RU] : Jr[6 + R0];
              w = valid ? fmaf(j2, u2, fmaf(j1, u1, j0 * u0)) : 0.f;
            } else {
              const float u0 = su[off0], u1 = su[L::U_SLAB + off0], u2 = su[2 * L::U_SLAB + off0];
              w = fmaf(Jr[6 + R0], u2, fmaf(Jr[3 + R0], u1, Jr[R0] * u0));
            }
          }
          split_tf32(w, ahi[kt][2 * c + h], alo[kt][2 * c + h]);
        }
      }
    }
    __syncwarp();
    if (nxt < nchunks) div32_issue(s, bar, &maps, Jg, ug, nxt, E, tma, lane);
    const unsigned tk = wq.ticket(lane);

    float acc[L::NT][4], corr[L::NT][4];
#pragma unroll
    for (int nt = 0; nt < L::NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) { acc[nt][q] = 0.f; corr[nt][q] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
      for (int nt = 0; nt < L::NT; ++nt) {
        mma_3xtf32(acc[nt], corr[nt], ahi[kt], alo[kt], sB[(kt * L::NT + nt) * 32 + lane]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < L::NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[nt][q] += corr[nt][q];
    // ---- stage [16][35], one TMA store ----
    const long long e0 = cur * kCH;
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < L::NT; ++nt) {
      const int i = 8 * nt + 2 * t;
      if (i < 35) { stage[g * 35 + i] = acc[nt][0]; stage[(g + 8) * 35 + i] = acc[nt][2]; }
      if (i + 1 < 35) { stage[g * 35 + i + 1] = acc[nt][1]; stage[(g + 8) * 35 + i + 1] = acc[nt][3]; }
    }
    fence_proxy_async();
    __syncwarp();
    if (tma) {
      if (lane == 0) { tma_store_2d(&maps.out, stage, 0, (int)(cur * (kCH / 4))); tma_store_commit(); }
    } else {
      flush_plain32(outg + e0 * 35, stage, e0, E, lane);
    }
    cur = nxt;
    nxt = wq.resolve(tk);
  }
  if (lane == 0) tma_store_wait_all();
}

// ================================================================ GRAD =====
// T[e][(i,r)] = sum_j u[e,j] D[r,i,j]  (K = 35 -> 40, N = 105 -> 14 tiles with the triple-aligned
// column layout of the fp64 kernel);  out[x,e,i] = sum_r J[x,r,e] T[e][(i,r)]
struct Grad32 {
  static constexpr int KT = 5, NTILE = 14;
  static constexpr int B_BYTES = 2 * NTILE * KT * 32 * 8;
  static constexpr int U_SLAB = kCH * 35;
  static constexpr int J_OFF = align128(U_SLAB * 4) / 4;
  static constexpr int SLOT_BYTES_TX = (U_SLAB + 9 * kCH) * 4;
  static constexpr int SLOT_BYTES = align128(J_OFF * 4 + 9 * kCH * 4);
  static constexpr int STAGE_BYTES = align128(3 * OUT_BLOCK32 * 4);
};

__device__ __forceinline__ void grad32_issue(float* s, uint64_t* bar, const OpMaps* maps, const float* __restrict__ Jg,
                                             const float* __restrict__ ug, long long chunk, long long E, bool tma, int lane) {
  using L = Grad32;
  const long long e0 = chunk * kCH;
  if (tma) {
    if (elect_one()) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, L::SLOT_BYTES_TX);
      tma_load_2d(s, &maps->in, 0, (int)(chunk * (kCH / 4)), bar);
      tma_load_2d(s + L::J_OFF, &maps->jac, (int)e0, 0, bar);
    }
  } else {
    const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
    for (int k = lane; k < L::U_SLAB; k += 32) s[k] = (k < ne * 35) ? ug[e0 * 35 + k] : 0.f;
    for (int k = lane; k < 9 * kCH; k += 32) {
      const int xr = k / kCH, el = k - xr * kCH;
      s[L::J_OFF + k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : 0.f;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  }
}

template <int T0, int NTG>
__device__ __forceinline__ void grad32_group(const uint4* __restrict__ sB,
                                             const uint32_t (&ahi)[Grad32::KT][4], const uint32_t (&alo)[Grad32::KT][4],
                                             const float (&Jr)[2][9], float* stage, int g, int t, int lane) {
  using L = Grad32;
  float acc[NTG][4], corr[NTG][4];
#pragma unroll
  for (int jt = 0; jt < NTG; ++jt)
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[jt][q] = 0.f; corr[jt][q] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
    for (int jt = 0; jt < NTG; ++jt) {
      mma_3xtf32(acc[jt], corr[jt], ahi[kt], alo[kt], sB[((T0 + jt) * L::KT + kt) * 32 + lane]);
    }
  }
#pragma unroll
  for (int jt = 0; jt < NTG; ++jt)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[jt][q] += corr[jt][q];
  if (T0 == 0) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  constexpr int V0 = 2 * T0;
  static_assert(V0 % 3 == 0, "groups must start on a triple boundary");
  constexpr int NTRI = (2 * NTG) / 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {                        // rows g and g + 8
    float* o = stage + (g + 8 * h) * 35 + 9 * t + V0 / 3;
#pragma unroll
    for (int q = 0; q < NTRI; ++q) {
      // value v = 3q + r of this group sits in tile v >> 1, column parity v & 1 (acc index 2h + parity)
      const float T0v = acc[(3 * q) >> 1][2 * h + ((3 * q) & 1)];
      const float T1v = acc[(3 * q + 1) >> 1][2 * h + ((3 * q + 1) & 1)];
      const float T2v = acc[(3 * q + 2) >> 1][2 * h + ((3 * q + 2) & 1)];
      if (V0 / 3 + q < 8 || t < 3) {
#pragma unroll
        for (int x = 0; x < 3; ++x)
          o[x * OUT_BLOCK32 + q] = fmaf(Jr[h][3 * x + 2], T2v, fmaf(Jr[h][3 * x + 1], T1v, Jr[h][3 * x] * T0v));
      }
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_grad_tf32(const __grid_constant__ OpMaps maps, const float* __restrict__ Jg, const float* __restrict__ Dg,
            const float* __restrict__ ug, float* __restrict__ outg, long long E, int flags) {
  using L = Grad32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* sB = reinterpret_cast<uint4*>(smem_raw);
  unsigned char* slots = smem_raw + L::B_BYTES;
  unsigned char* stages = slots + (size_t)NW * L::SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)NW * L::STAGE_BYTES);
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], 1);
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();

  float* s = reinterpret_cast<float*>(slots + (size_t)warp * L::SLOT_BYTES);
  float* stage = reinterpret_cast<float*>(stages + (size_t)warp * L::STAGE_BYTES);
  uint64_t* bar = &bars[warp];
  const float* sJ = s + L::J_OFF;
  const long long nchunks = (E + kCH - 1) / kCH;
  const WorkQueue wq{work_ctr, nchunks};
  const int g = lane >> 2, t = lane & 3;
  const bool tma = flags & kFlagTma;

  long long cur = wq.take(lane), nxt = wq.take(lane);
  if (cur < nchunks) grad32_issue(s, bar, &maps, Jg, ug, cur, E, tma, lane);
  // operator tables are staged while the first TMA loads are in flight
  // fragment (tile, kt): column c = g holds value v = 2 tile + (c & 1) of lane c >> 1: (dof 9 (c>>1) + v/3, r = v%3)
  fill_b_table(sB, L::NTILE * L::KT, [&](int frag, int ln, int half) {
    const int tile = frag / L::KT, kt = frag - tile * L::KT;
    const int c = ln >> 2, t = ln & 3, j = 8 * kt + t + 4 * half;
    const int v = 2 * tile + (c & 1), i = 9 * (c >> 1) + v / 3, r = v % 3;
    return (v < 27 && i < 35 && j < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.f;
  });
  __syncthreads();
  for (uint32_t n = 0; cur < nchunks; ++n) {
    mbar_wait(bar, n & 1u);
    uint32_t ahi[L::KT][4], alo[L::KT][4];
    float Jr[2][9];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int el = g + 8 * h;
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jr[h][xr] = sJ[xr * kCH + el];
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = 8 * kt + 4 * c + t;
          float v = s[el * 35 + (8 * kt + 4 * c < 35 ? j : 0)];
          if (8 * kt + 4 * c + 3 >= 35 && j >= 35) v = 0.f;     // j = 35 .. 39 is padding
          if (8 * kt + 4 * c >= 35) v = 0.f;
          split_tf32(v, ahi[kt][2 * c + h], alo[kt][2 * c + h]);
        }
      }
    }
    __syncwarp();
    if (nxt < nchunks) grad32_issue(s, bar, &maps, Jg, ug, nxt, E, tma, lane);
    const unsigned tk = wq.ticket(lane);

    const long long e0 = cur * kCH;
    grad32_group<0, 3>(sB, ahi, alo, Jr, stage, g, t, lane);
    grad32_group<3, 3>(sB, ahi, alo, Jr, stage, g, t, lane);
    grad32_group<6, 3>(sB, ahi, alo, Jr, stage, g, t, lane);
    grad32_group<9, 3>(sB, ahi, alo, Jr, stage, g, t, lane);
    grad32_group<12, 2>(sB, ahi, alo, Jr, stage, g, t, lane);
    fence_proxy_async();
    __syncwarp();
    if (tma) {
      if (lane == 0) { tma_store_3d(&maps.out, stage, 0, (int)(cur * (kCH / 4)), 0); tma_store_commit(); }
    } else {
#pragma unroll
      for (int x = 0; x < 3; ++x)
        flush_plain32(outg + ((long long)x * E + e0) * 35, stage + x * OUT_BLOCK32, e0, E, lane);
    }
    cur = nxt;
    nxt = wq.resolve(tk);
  }
  if (lane == 0) tma_store_wait_all();
}

// ================================================================ LIFT =====
// out_k[e,i] = sum_k Op(f,i,j) Jf(e,f) v_k[f,e,j],  k = 15 f + j (60 -> 64)
struct Lift32 {
  static constexpr int KT = 8, NT = 5;
  static constexpr int B_BYTES = 2 * KT * NT * 32 * 8;
  static constexpr int V_SLAB = kCH * 15;                               // floats per face
  static constexpr int J_OFF = 4 * V_SLAB;                              // 3840 floats = 15360 B (128-B aligned)
  static constexpr int SLOT_BYTES_TX = (4 * V_SLAB + 4 * kCH) * 4;
  static constexpr int SLOT_BYTES = align128(SLOT_BYTES_TX);
  static constexpr int STAGE_BYTES = align128(OUT_BLOCK32 * 4);
};

template <bool FE>
__device__ __forceinline__ void lift32_issue(float* s, uint64_t* bar, const CUtensorMap* map_v, const CUtensorMap* map_j,
                                             const float* __restrict__ Jg, const float* __restrict__ vg,
                                             long long chunk, long long E, bool tma, int lane) {
  using L = Lift32;
  const long long e0 = chunk * kCH;
  if (tma) {
    if (elect_one()) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, L::SLOT_BYTES_TX);
      tma_load_3d(s, map_v, 0, (int)(chunk * (kCH / 4)), 0, bar);
      if (FE) tma_load_2d(s + L::J_OFF, map_j, (int)e0, 0, bar);
      else    tma_load_2d(s + L::J_OFF, map_j, 0, (int)(chunk * (kCH / 4)), bar);
    }
  } else {
    const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
    for (int f = 0; f < 4; ++f)
      for (int k = lane; k < L::V_SLAB; k += 32)
        s[f * L::V_SLAB + k] = (k < ne * 15) ? vg[((long long)f * E + e0) * 15 + k] : 0.f;
    for (int k = lane; k < 4 * kCH; k += 32) {
      float v = 0.f;
      if (FE) { const int f = k / kCH, el = k - f * kCH; if (el < ne) v = Jg[(long long)f * E + e0 + el]; }
      else    { const int el = k / 4; if (el < ne) v = Jg[e0 * 4 + k]; }
      s[L::J_OFF + k] = v;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  }
}

template <int NW, bool FE>
__global__ void __launch_bounds__(NW * 32, 1)
k_lift_tf32(const __grid_constant__ LiftMaps maps, const float* __restrict__ Jg, const float* __restrict__ Og,
            const __grid_constant__ OpmatRows rows, int nrows, long long E, int flags) {
  using L = Lift32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* sB = reinterpret_cast<uint4*>(smem_raw);
  unsigned char* slots = smem_raw + L::B_BYTES;
  unsigned char* stages = slots + (size_t)NW * L::SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)NW * L::STAGE_BYTES);
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], 1);
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();

  float* s = reinterpret_cast<float*>(slots + (size_t)warp * L::SLOT_BYTES);
  float* stage = reinterpret_cast<float*>(stages + (size_t)warp * L::STAGE_BYTES);
  uint64_t* bar = &bars[warp];
  const float* sJ = s + L::J_OFF;
  const long long nchunks = (E + kCH - 1) / kCH;
  const WorkQueue wq{work_ctr, nchunks};
  const int g = lane >> 2, t = lane & 3;
  const bool tma = flags & kFlagTma;

  long long cur = wq.take(lane), nxt = wq.take(lane);
  int fld = 0;
  if (cur < nchunks)
    lift32_issue<FE>(s, bar, &maps.in[0], &maps.jac, Jg, static_cast<const float*>(rows.field[0]), cur, E, tma, lane);
  // operator tables are staged while the first TMA loads are in flight
  fill_b_table(sB, L::KT * L::NT, [&](int frag, int ln, int half) {
    const int kt = frag / L::NT, nt = frag - kt * L::NT;
    const int g = ln >> 2, t = ln & 3, k = 8 * kt + t + 4 * half, i = 8 * nt + g;
    const int f = k / 15, j = k - 15 * f;
    if (k >= 60 || i >= 35) return 0.f;
    return FE ? Og[(i * 4 + f) * 15 + j] : Og[(f * 35 + i) * 15 + j];
  });
  __syncthreads();
  for (uint32_t n = 0; cur < nchunks; ++n) {
    mbar_wait(bar, n & 1u);
    uint32_t ahi[L::KT][4], alo[L::KT][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int el = g + 8 * h;
      float Jf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) Jf[f] = FE ? sJ[f * kCH + el] : sJ[el * 4 + f];
      const float* sv = s + el * 15 + t;
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // k = 8kt + 4c + t = 15 f + j
          const int K0 = 8 * kt + 4 * c, F0 = K0 / 15, THR = 15 * (F0 + 1) - K0;
          float w = 0.f;
          if (K0 < 60) {
            const int off0 = K0 + (L::V_SLAB - 15) * F0;       // f*V_SLAB + j for t = 0, minus t
            if (THR < 4) {
              const bool up = t >= THR;
              const bool valid = !up || F0 + 1 < 4;
              const int FU = F0 + 1 < 4 ? F0 + 1 : 3;
              const float v = sv[valid ? (up ? off0 + (L::V_SLAB - 15) : off0) : 0];
              w = valid ? (up ? Jf[FU] : Jf[F0]) * v : 0.f;
            } else {
              w = Jf[F0] * sv[off0];
            }
          }
          split_tf32(w, ahi[kt][2 * c + h], alo[kt][2 * c + h]);
        }
      }
    }
    __syncwarp();
    const bool advance = fld + 1 == nrows;
    const int nfld = advance ? 0 : fld + 1;
    const long long nchunk = advance ? nxt : cur;
    if (nchunk < nchunks)
      lift32_issue<FE>(s, bar, &maps.in[nfld], &maps.jac, Jg, static_cast<const float*>(rows.field[nfld]), nchunk, E, tma, lane);
    unsigned tk = 0;
    if (advance) tk = wq.ticket(lane);

    float acc[L::NT][4], corr[L::NT][4];
#pragma unroll
    for (int nt = 0; nt < L::NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) { acc[nt][q] = 0.f; corr[nt][q] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
      for (int nt = 0; nt < L::NT; ++nt) {
        mma_3xtf32(acc[nt], corr[nt], ahi[kt], alo[kt], sB[(kt * L::NT + nt) * 32 + lane]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < L::NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[nt][q] += corr[nt][q];
    const long long e0 = cur * kCH;
    float* outg = static_cast<float*>(rows.out[fld]);
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < L::NT; ++nt) {
      const int i = 8 * nt + 2 * t;
      if (i < 35) { stage[g * 35 + i] = acc[nt][0]; stage[(g + 8) * 35 + i] = acc[nt][2]; }
      if (i + 1 < 35) { stage[g * 35 + i + 1] = acc[nt][1]; stage[(g + 8) * 35 + i + 1] = acc[nt][3]; }
    }
    fence_proxy_async();
    __syncwarp();
    if (tma) {
      if (lane == 0) { tma_store_2d(&maps.out[fld], stage, 0, (int)(cur * (kCH / 4))); tma_store_commit(); }
    } else {
      flush_plain32(outg + e0 * 35, stage, e0, E, lane);
    }
    if (advance) { cur = nxt; nxt = wq.resolve(tk); }
    fld = nfld;
  }
  if (lane == 0) tma_store_wait_all();
}

// ------------------------------------------------------------ launchers ----
static bool make_map32(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides, const cuuint32_t* box) {
  return make_map_typed(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides, box);
}
// (E, W) fp32 rows viewed as (E/4, 4W): box = 4 quads = one 16-element chunk
static bool map32_rows(CUtensorMap* tm, const void* base, long long E, int W) {
  const cuuint64_t dims[2] = {(cuuint64_t)(4 * W), (cuuint64_t)(E / 4)};
  const cuuint64_t strides[1] = {(cuuint64_t)(16 * W)};
  const cuuint32_t box[2] = {(cuuint32_t)(4 * W), (cuuint32_t)(kCH / 4)};
  return make_map32(tm, base, 2, dims, strides, box);
}
static bool map32_slabs(CUtensorMap* tm, const void* base, long long E, int W, int S) {
  const cuuint64_t dims[3] = {(cuuint64_t)(4 * W), (cuuint64_t)(E / 4), (cuuint64_t)S};
  const cuuint64_t strides[2] = {(cuuint64_t)(16 * W), (cuuint64_t)E * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)(4 * W), (cuuint32_t)(kCH / 4), (cuuint32_t)S};
  return make_map32(tm, base, 3, dims, strides, box);
}
static bool map32_erows(CUtensorMap* tm, const void* base, long long E, int R) {
  const cuuint64_t dims[2] = {(cuuint64_t)E, (cuuint64_t)R};
  const cuuint64_t strides[1] = {(cuuint64_t)E * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kCH, (cuuint32_t)R};
  return make_map32(tm, base, 2, dims, strides, box);
}

template <int NW>
static int launch_tf32_nw(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                          long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  const long long nchunks = (E + kCH - 1) / kCH;
  const float* J = static_cast<const float*>(jac);
  const float* O = static_cast<const float*>(op);
  constexpr int threads = NW * 32;
  bool tma = (E % 4 == 0) && E < (1LL << 31) - kCH && aligned16(jac);
  for (int r = 0; r < nrows; ++r) tma = tma && aligned16(rows.field[r]) && aligned16(rows.out[r]);
  if (cfg && (cfg->reserved[0] & 1)) tma = false;
  const long long need = (nchunks + NW - 1) / NW;
  const unsigned grid = (unsigned)(di.sms < need ? di.sms : need);

  if (kind == FNSM_OP_DIV || kind == FNSM_OP_GRAD) {
    const bool is_div = kind == FNSM_OP_DIV;
    const size_t smem = (is_div ? Div32::B_BYTES + (size_t)NW * (Div32::SLOT_BYTES + Div32::STAGE_BYTES)
                                : Grad32::B_BYTES + (size_t)NW * (Grad32::SLOT_BYTES + Grad32::STAGE_BYTES)) +
                        8 * (size_t)NW + 8;
    if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
    for (int r = 0; r < nrows; ++r) {
      const float* u = static_cast<const float*>(rows.field[r]);
      float* out = static_cast<float*>(rows.out[r]);
      OpMaps maps;
      bool ok = tma && map32_erows(&maps.jac, J, E, 9);
      if (is_div) ok = ok && map32_slabs(&maps.in, u, E, 35, 3) && map32_rows(&maps.out, out, E, 35);
      else        ok = ok && map32_rows(&maps.in, u, E, 35) && map32_slabs(&maps.out, out, E, 35, 3);
      const int flags = ok ? kFlagTma : 0;
      if (is_div) {
        if (int rc = set_smem(k_div_tf32<NW>, smem)) return rc;
        k_div_tf32<NW><<<grid, threads, smem, st>>>(maps, J, O, u, out, E, flags);
      } else {
        if (int rc = set_smem(k_grad_tf32<NW>, smem)) return rc;
        k_grad_tf32<NW><<<grid, threads, smem, st>>>(maps, J, O, u, out, E, flags);
      }
      if (int rc = post_launch()) return rc;
    }
    return FNSM_OK;
  }
  const size_t smem = Lift32::B_BYTES + (size_t)NW * (Lift32::SLOT_BYTES + Lift32::STAGE_BYTES) + 8 * (size_t)NW + 8;
  if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  LiftMaps maps;
  bool ok = tma && (kind == FNSM_OP_LIFT_FE ? map32_erows(&maps.jac, J, E, 4) : map32_rows(&maps.jac, J, E, 4));
  for (int r = 0; r < nrows && ok; ++r)
    ok = map32_slabs(&maps.in[r], rows.field[r], E, 15, 4) && map32_rows(&maps.out[r], rows.out[r], E, 35);
  const int flags = ok ? kFlagTma : 0;
  if (kind == FNSM_OP_LIFT_FE) {
    if (int rc = set_smem(k_lift_tf32<NW, true>, smem)) return rc;
    k_lift_tf32<NW, true><<<grid, threads, smem, st>>>(maps, J, O, rows, nrows, E, flags);
  } else {
    if (int rc = set_smem(k_lift_tf32<NW, false>, smem)) return rc;
    k_lift_tf32<NW, false><<<grid, threads, smem, st>>>(maps, J, O, rows, nrows, E, flags);
  }
  return post_launch();
}

static int launch_tf32(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                       long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  if (cfg && cfg->ctas_per_sm > 1) return FNSM_E_BAD_CONFIG;
  if (cfg && (cfg->stages < 0 || cfg->stages > 1)) return FNSM_E_BAD_CONFIG;
  // round-1 sweep: 16 warps for grad / lift, 12 for div (its 112 A-fragment registers spill at 16)
  const int dflt = kind == FNSM_OP_DIV ? 384 : 512;
  const int threads = (cfg && cfg->threads != 0) ? cfg->threads : dflt;
  switch (threads) {
    case 128: return launch_tf32_nw<4>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 256: return launch_tf32_nw<8>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 384: return launch_tf32_nw<12>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 512: return launch_tf32_nw<16>(kind, jac, op, rows, nrows, E, cfg, di, st);
    default: return FNSM_E_BAD_CONFIG;
  }
}

}  // namespace fnsm
