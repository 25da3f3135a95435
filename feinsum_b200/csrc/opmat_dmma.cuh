// Variant 1 ("dmma") of the DG operator kernels for fp64, p = 4 tets.
//
// Every DG einsum is (tiny per-element scaling) o (constant matrix x element
// vector) -- SURVEY.md Appendix C.  Here the constant-matrix part runs on the
// FP64 tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no
// FP64 kind), with
//     M = 8 elements,   N = 8 output dofs (or (r, dof) pairs),   K = 4 contracted dofs
//   * B fragments (the operator, zero padded to multiples of 8 x 4) are laid
//     out once per CTA in shared memory in fragment order [k-tile][n-tile][lane]
//     -> conflict-free LDS.64, shared by the ME = 2 element tiles a warp owns;
//   * A fragments (per-element data) are produced on the fly from the element
//     slot: div folds the Jacobian (w = sum_x J[x,r,e] u[x,e,j]), lift scales
//     by the face Jacobian, grad reads u directly and applies J to the
//     accumulator fragments afterwards -- the hoisting the reference's
//     transforms perform (tuning/impls/xre_rij_ej_to_xei.py:104-117,
//     xre_rij_xej_to_ei_v6.py:212-248, ifj_fe_fej_to_ei_v3.py:94-105);
//   * one persistent CTA per SM: NW consumer warps + 1 producer warp.  The
//     producer streams 16-element chunks into a ring of shared-memory slots
//     with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); a
//     consumer warp owns one chunk at a time, so there is no CTA-wide barrier
//     in steady state;
//   * elements are permuted inside a 16-chunk (el = 4*(g&3) + (g>>2) + 2m) so
//     that the stride-35 rows read by a half-warp fall into distinct banks.
//
// Unaligned inputs (odd E, tail chunk) take a plain-load path in the producer.
#pragma once
#include "common.cuh"
#include "opmat_simt.cuh"

namespace fnsm {

// ----------------------------------------------------------------- PTX -----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n"
      "DONE_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA unit)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ------------------------------------------------------------ geometry -----
constexpr int kME = 2;            // element tiles (of 8) per warp chunk
constexpr int kCH = 8 * kME;      // elements per chunk / ring slot
constexpr int kMaxWarps = 12;

struct DmmaShape {
  // p = 4 tets
  static constexpr int NI = 35, NJ = 35, ND = 3, NF = 4, NFJ = 15;
  static constexpr int NT = 5;                 // n-tiles of 8 covering NI (padded to 40)
};

__device__ __forceinline__ int chunk_el(int g, int m) { return 4 * (g & 3) + (g >> 2) + 2 * m; }

// ================================================================= DIV =====
// out[e,i] = sum_{r,j} D[r,i,j] * (sum_x J[x,r,e] u[x,e,j])
// k-tiles ordered (jq, r): kt = 3*jq + r, k-in-tile t <-> j = 4*jq + t
struct DivLayout {
  static constexpr int KT = 27;
  static constexpr int B_DOUBLES = KT * DmmaShape::NT * 32;          // 4320
  static constexpr int U_SLAB = kCH * 35;                            // doubles per x
  static constexpr int SLOT_DOUBLES = 3 * U_SLAB + 9 * kCH;          // 1824 -> 14592 B
  static constexpr uint32_t SLOT_BYTES = SLOT_DOUBLES * 8;
};

template <int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
k_div_dmma(const double* __restrict__ Jg, const double* __restrict__ Dg,
           const double* __restrict__ ug, double* __restrict__ outg,
           long long E, int nslots, int tma_ok) {
  using L = DivLayout;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* ring = sB + L::B_DOUBLES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)nslots * L::SLOT_DOUBLES);
  uint64_t* empty = full + nslots;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operator fragments: sB[(kt*NT + nt)*32 + lane] = D[r][8nt+g][4jq+t]
  for (int idx = threadIdx.x; idx < L::B_DOUBLES; idx += blockDim.x) {
    const int ln = idx & 31, nt = (idx >> 5) % DmmaShape::NT, kt = (idx >> 5) / DmmaShape::NT;
    const int g = ln >> 2, t = ln & 3, jq = kt / 3, r = kt - 3 * jq;
    const int i = 8 * nt + g, j = 4 * jq + t;
    sB[idx] = (i < 35 && j < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.0;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  const long long nchunks = (E + kCH - 1) / kCH;
  const long long G = gridDim.x;
  // number of chunks this CTA processes: b, b+G, ...
  const long long nq = (nchunks > (long long)blockIdx.x) ? (nchunks - blockIdx.x + G - 1) / G : 0;

  if (warp == NW) {
    // ------------------------------------------------------ producer ------
    for (long long q = 0; q < nq; ++q) {
      const int slot = (int)(q % nslots);
      const long long use = q / nslots;
      if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
      const long long e0 = (blockIdx.x + q * G) * kCH;
      double* s = ring + (size_t)slot * L::SLOT_DOUBLES;
      if (tma_ok && e0 + kCH <= E) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[slot], L::SLOT_BYTES);
          for (int x = 0; x < 3; ++x)
            tma_load_1d(s + x * L::U_SLAB, ug + ((long long)x * E + e0) * 35, L::U_SLAB * 8, &full[slot]);
          for (int xr = 0; xr < 9; ++xr)
            tma_load_1d(s + 3 * L::U_SLAB + xr * kCH, Jg + (long long)xr * E + e0, kCH * 8, &full[slot]);
        }
      } else {
        const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
        for (int x = 0; x < 3; ++x)
          for (int k = lane; k < L::U_SLAB; k += 32)
            s[x * L::U_SLAB + k] = (k < ne * 35) ? ug[((long long)x * E + e0) * 35 + k] : 0.0;
        for (int k = lane; k < 9 * kCH; k += 32) {
          const int xr = k / kCH, el = k - xr * kCH;
          s[3 * L::U_SLAB + k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[slot]);
      }
    }
    return;
  }

  // -------------------------------------------------------- consumers -----
  const int g = lane >> 2, t = lane & 3;
  for (long long q = warp; q < nq; q += NW) {
    const int slot = (int)(q % nslots);
    const long long use = q / nslots;
    mbar_wait(&full[slot], (uint32_t)(use & 1));
    const double* s = ring + (size_t)slot * L::SLOT_DOUBLES;
    const double* sJ = s + 3 * L::U_SLAB;

    double acc[kME][DmmaShape::NT][2];
    double Jr[kME][9];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jr[m][xr] = sJ[xr * kCH + el];
#pragma unroll
      for (int nt = 0; nt < DmmaShape::NT; ++nt) { acc[m][nt][0] = 0.0; acc[m][nt][1] = 0.0; }
    }
#pragma unroll
    for (int jq = 0; jq < 9; ++jq) {
      double ux[kME][3];
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        const int el = chunk_el(g, m);
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          double v = s[x * L::U_SLAB + el * 35 + 4 * jq + t];
          if (jq == 8 && t == 3) v = 0.0;           // j = 35 is padding
          ux[m][x] = v;
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        double a[kME];
#pragma unroll
        for (int m = 0; m < kME; ++m)
          a[m] = fma(Jr[m][6 + r], ux[m][2], fma(Jr[m][3 + r], ux[m][1], Jr[m][r] * ux[m][0]));
        const double* bp = sB + ((3 * jq + r) * DmmaShape::NT) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < DmmaShape::NT; ++nt) {
          const double b = bp[nt * 32];
#pragma unroll
          for (int m = 0; m < kME; ++m) dmma884(acc[m][nt], a[m], b);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);     // slot data no longer needed

    const long long e0 = (blockIdx.x + q * G) * kCH;
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const long long e = e0 + chunk_el(g, m);
      if (e < E) {
        double* o = outg + e * 35;
#pragma unroll
        for (int nt = 0; nt < DmmaShape::NT; ++nt) {
          const int i = 8 * nt + 2 * t;
          if (i < 35) stg_stream(o + i, acc[m][nt][0]);
          if (i + 1 < 35) stg_stream(o + i + 1, acc[m][nt][1]);
        }
      }
    }
  }
}

// ================================================================ GRAD =====
// T[r][e][i] = sum_j D[r,i,j] u[e,j];  out[x,e,i] = sum_r J[x,r,e] T[r][e][i]
// n-tiles ordered (it, r); k-tiles kt <-> j = 4*kt + t (9 tiles)
struct GradLayout {
  static constexpr int KT = 9;
  static constexpr int B_DOUBLES = DmmaShape::NT * KT * 3 * 32;       // 4320
  static constexpr int U_SLAB = kCH * 35;
  static constexpr int SLOT_DOUBLES = U_SLAB + 9 * kCH;               // 704 -> 5632 B
  static constexpr uint32_t SLOT_BYTES = SLOT_DOUBLES * 8;
};

template <int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
k_grad_dmma(const double* __restrict__ Jg, const double* __restrict__ Dg,
            const double* __restrict__ ug, double* __restrict__ outg,
            long long E, int nslots, int tma_ok) {
  using L = GradLayout;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* ring = sB + L::B_DOUBLES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)nslots * L::SLOT_DOUBLES);
  uint64_t* empty = full + nslots;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // sB[((it*KT + kt)*3 + r)*32 + lane] = D[r][8it+g][4kt+t]
  for (int idx = threadIdx.x; idx < L::B_DOUBLES; idx += blockDim.x) {
    const int ln = idx & 31, r = (idx >> 5) % 3, kt = ((idx >> 5) / 3) % L::KT, it = (idx >> 5) / (3 * L::KT);
    const int g = ln >> 2, t = ln & 3;
    const int i = 8 * it + g, j = 4 * kt + t;
    sB[idx] = (i < 35 && j < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.0;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  const long long nchunks = (E + kCH - 1) / kCH;
  const long long G = gridDim.x;
  const long long nq = (nchunks > (long long)blockIdx.x) ? (nchunks - blockIdx.x + G - 1) / G : 0;

  if (warp == NW) {
    for (long long q = 0; q < nq; ++q) {
      const int slot = (int)(q % nslots);
      const long long use = q / nslots;
      if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
      const long long e0 = (blockIdx.x + q * G) * kCH;
      double* s = ring + (size_t)slot * L::SLOT_DOUBLES;
      if (tma_ok && e0 + kCH <= E) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[slot], L::SLOT_BYTES);
          tma_load_1d(s, ug + e0 * 35, L::U_SLAB * 8, &full[slot]);
          for (int xr = 0; xr < 9; ++xr)
            tma_load_1d(s + L::U_SLAB + xr * kCH, Jg + (long long)xr * E + e0, kCH * 8, &full[slot]);
        }
      } else {
        const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
        for (int k = lane; k < L::U_SLAB; k += 32) s[k] = (k < ne * 35) ? ug[e0 * 35 + k] : 0.0;
        for (int k = lane; k < 9 * kCH; k += 32) {
          const int xr = k / kCH, el = k - xr * kCH;
          s[L::U_SLAB + k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[slot]);
      }
    }
    return;
  }

  const int g = lane >> 2, t = lane & 3;
  for (long long q = warp; q < nq; q += NW) {
    const int slot = (int)(q % nslots);
    const long long use = q / nslots;
    mbar_wait(&full[slot], (uint32_t)(use & 1));
    const double* s = ring + (size_t)slot * L::SLOT_DOUBLES;
    const double* sJ = s + L::U_SLAB;

    double a[kME][L::KT];
    double Jr[kME][9];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
        double v = s[el * 35 + 4 * kt + t];
        if (kt == 8 && t == 3) v = 0.0;
        a[m][kt] = v;
      }
    }
    // J is needed for the accumulator columns n = 2t, 2t+1 (elements of row-group
    // 2t / 2t+1 of the C fragment), not for row g: C[row g][col] has row = element.
    // Here M = element (row g), so J[x][r][el(g)] scales this thread's two values.
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jr[m][xr] = sJ[xr * kCH + el];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);     // everything is in registers now

    const long long e0 = (blockIdx.x + q * G) * kCH;
#pragma unroll 1
    for (int it = 0; it < DmmaShape::NT; ++it) {
      double acc[kME][3][2];
#pragma unroll
      for (int m = 0; m < kME; ++m)
#pragma unroll
        for (int r = 0; r < 3; ++r) { acc[m][r][0] = 0.0; acc[m][r][1] = 0.0; }
      const double* bp = sB + (size_t)it * L::KT * 3 * 32 + lane;
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double b = bp[(kt * 3 + r) * 32];
#pragma unroll
          for (int m = 0; m < kME; ++m) dmma884(acc[m][r], a[m][kt], b);
        }
      }
      const int i = 8 * it + 2 * t;
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        const long long e = e0 + chunk_el(g, m);
        if (e < E) {
#pragma unroll
          for (int x = 0; x < 3; ++x) {
            const double o0 = fma(Jr[m][3 * x + 2], acc[m][2][0], fma(Jr[m][3 * x + 1], acc[m][1][0], Jr[m][3 * x] * acc[m][0][0]));
            const double o1 = fma(Jr[m][3 * x + 2], acc[m][2][1], fma(Jr[m][3 * x + 1], acc[m][1][1], Jr[m][3 * x] * acc[m][0][1]));
            double* o = outg + ((long long)x * E + e) * 35;
            if (i < 35) stg_stream(o + i, o0);
            if (i + 1 < 35) stg_stream(o + i + 1, o1);
          }
        }
      }
    }
  }
}

// ================================================================ LIFT =====
// out_k[e,i] = sum_{f,j} Op(f,i,j) * Jf(e,f) * v_k[f,e,j];  K = (f,j) = 60 = 15 k-tiles
// work item = (chunk, field): the ring streams one field of one chunk per slot
struct LiftLayout {
  static constexpr int KT = 15;
  static constexpr int B_DOUBLES = KT * DmmaShape::NT * 32;           // 2400
  static constexpr int V_SLAB = kCH * 15;                             // doubles per face
  static constexpr int SLOT_DOUBLES = 4 * V_SLAB + 4 * kCH;           // 1024 -> 8192 B
  static constexpr uint32_t SLOT_BYTES = SLOT_DOUBLES * 8;
};

template <int NW, bool FE>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
k_lift_dmma(const double* __restrict__ Jg, const double* __restrict__ Og, OpmatRows rows, int nrows,
            long long E, int nslots, int tma_ok) {
  using L = LiftLayout;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* ring = sB + L::B_DOUBLES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)nslots * L::SLOT_DOUBLES);
  uint64_t* empty = full + nslots;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // sB[(kt*NT + nt)*32 + lane] = Op(f, 8nt+g, j),  k = 4kt+t = 15 f + j
  for (int idx = threadIdx.x; idx < L::B_DOUBLES; idx += blockDim.x) {
    const int ln = idx & 31, nt = (idx >> 5) % DmmaShape::NT, kt = (idx >> 5) / DmmaShape::NT;
    const int g = ln >> 2, t = ln & 3, k = 4 * kt + t, f = k / 15, j = k - 15 * f;
    const int i = 8 * nt + g;
    double v = 0.0;
    if (i < 35) v = FE ? Og[(i * 4 + f) * 15 + j] : Og[(f * 35 + i) * 15 + j];
    sB[idx] = v;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  const long long nchunks = (E + kCH - 1) / kCH;
  const long long nitems = nchunks * nrows;      // item = chunk * nrows + field
  const long long G = gridDim.x;
  // CTA b owns chunks b, b+G, ...; all fields of a chunk are consecutive items
  const long long nqc = (nchunks > (long long)blockIdx.x) ? (nchunks - blockIdx.x + G - 1) / G : 0;
  const long long nq = nqc * nrows;
  (void)nitems;

  if (warp == NW) {
    for (long long q = 0; q < nq; ++q) {
      const int slot = (int)(q % nslots);
      const long long use = q / nslots;
      if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
      const long long qc = q / nrows;
      const int fld = (int)(q - qc * nrows);
      const long long e0 = (blockIdx.x + qc * G) * kCH;
      const double* vg = static_cast<const double*>(rows.field[fld]);
      double* s = ring + (size_t)slot * L::SLOT_DOUBLES;
      const bool fast = tma_ok && e0 + kCH <= E;
      if (fast) {
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[slot], L::SLOT_BYTES);
          for (int f = 0; f < 4; ++f)
            tma_load_1d(s + f * L::V_SLAB, vg + ((long long)f * E + e0) * 15, L::V_SLAB * 8, &full[slot]);
          if (FE) {
            for (int f = 0; f < 4; ++f)   // Jface(f, e): 4 rows of kCH
              tma_load_1d(s + 4 * L::V_SLAB + f * kCH, Jg + (long long)f * E + e0, kCH * 8, &full[slot]);
          } else {                        // J(e, f): kCH*4 contiguous
            tma_load_1d(s + 4 * L::V_SLAB, Jg + e0 * 4, 4 * kCH * 8, &full[slot]);
          }
        }
      } else {
        const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
        for (int f = 0; f < 4; ++f)
          for (int k = lane; k < L::V_SLAB; k += 32)
            s[f * L::V_SLAB + k] = (k < ne * 15) ? vg[((long long)f * E + e0) * 15 + k] : 0.0;
        for (int k = lane; k < 4 * kCH; k += 32) {
          double v = 0.0;
          if (FE) { const int f = k / kCH, el = k - f * kCH; if (el < ne) v = Jg[(long long)f * E + e0 + el]; }
          else    { const int el = k / 4; if (el < ne) v = Jg[e0 * 4 + k]; }
          s[4 * L::V_SLAB + k] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[slot]);
      }
    }
    return;
  }

  const int g = lane >> 2, t = lane & 3;
  for (long long q = warp; q < nq; q += NW) {
    const int slot = (int)(q % nslots);
    const long long use = q / nslots;
    mbar_wait(&full[slot], (uint32_t)(use & 1));
    const double* s = ring + (size_t)slot * L::SLOT_DOUBLES;
    const double* sJ = s + 4 * L::V_SLAB;

    double acc[kME][DmmaShape::NT][2];
    double Jf[kME][4];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
#pragma unroll
      for (int f = 0; f < 4; ++f) Jf[m][f] = FE ? sJ[f * kCH + el] : sJ[el * 4 + f];
#pragma unroll
      for (int nt = 0; nt < DmmaShape::NT; ++nt) { acc[m][nt][0] = 0.0; acc[m][nt][1] = 0.0; }
    }
#pragma unroll
    for (int kt = 0; kt < L::KT; ++kt) {
      double a[kME];
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        const int el = chunk_el(g, m);
        // k = 4kt + t = 15 f + j ; f and j are not compile-time (t is a lane id)
        const int k = 4 * kt + t, f = k / 15, j = k - 15 * f;
        const double jf = (f == 0) ? Jf[m][0] : (f == 1) ? Jf[m][1] : (f == 2) ? Jf[m][2] : Jf[m][3];
        a[m] = jf * s[f * L::V_SLAB + el * 15 + j];
      }
      const double* bp = sB + (kt * DmmaShape::NT) * 32 + lane;
#pragma unroll
      for (int nt = 0; nt < DmmaShape::NT; ++nt) {
        const double b = bp[nt * 32];
#pragma unroll
        for (int m = 0; m < kME; ++m) dmma884(acc[m][nt], a[m], b);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);

    const long long qc = q / nrows;
    const int fld = (int)(q - qc * nrows);
    const long long e0 = (blockIdx.x + qc * G) * kCH;
    double* outg = static_cast<double*>(rows.out[fld]);
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const long long e = e0 + chunk_el(g, m);
      if (e < E) {
        double* o = outg + e * 35;
#pragma unroll
        for (int nt = 0; nt < DmmaShape::NT; ++nt) {
          const int i = 8 * nt + 2 * t;
          if (i < 35) stg_stream(o + i, acc[m][nt][0]);
          if (i + 1 < 35) stg_stream(o + i + 1, acc[m][nt][1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------ launchers ----
inline bool dmma_supported(int kind, int n_outer, int ni, int nj) {
  if (kind == FNSM_OP_GRAD || kind == FNSM_OP_DIV) return n_outer == 3 && ni == 35 && nj == 35;
  return n_outer == 4 && ni == 35 && nj == 15;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class K>
static int set_smem(K kernel, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return e == cudaSuccess ? FNSM_OK : (int)e;
}

static int launch_dmma(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                       int n_outer, int ni, int nj, long long E, const fnsm_cfg* cfg,
                       const DevInfo& di, cudaStream_t st) {
  (void)n_outer; (void)ni; (void)nj;
  constexpr int NW = 8;
  if (cfg && cfg->threads != 0 && cfg->threads != (NW + 1) * 32) return FNSM_E_BAD_CONFIG;
  int extra = (cfg && cfg->stages > 0) ? cfg->stages : 4;
  if (extra < 1 || extra > 24) return FNSM_E_BAD_CONFIG;
  int nslots = NW + extra;
  const long long nchunks = (E + kCH - 1) / kCH;
  long long grid = di.sms;   // one persistent CTA per SM (launch bound 1 CTA/SM)
  if (cfg && cfg->ctas_per_sm > 1) return FNSM_E_BAD_CONFIG;
  if (grid > nchunks) grid = nchunks;
  const double* J = static_cast<const double*>(jac);
  const double* O = static_cast<const double*>(op);
  const int threads = (NW + 1) * 32;
  int tma_ok = (E % 2 == 0) && aligned16(jac);
  for (int r = 0; r < nrows; ++r) tma_ok = tma_ok && aligned16(rows.field[r]);

  if (kind == FNSM_OP_DIV || kind == FNSM_OP_GRAD) {
    const bool is_div = kind == FNSM_OP_DIV;
    const size_t slot_d = is_div ? DivLayout::SLOT_DOUBLES : GradLayout::SLOT_DOUBLES;
    const size_t b_d = is_div ? DivLayout::B_DOUBLES : GradLayout::B_DOUBLES;
    size_t smem = 8 * (b_d + (size_t)nslots * slot_d) + 16 * (size_t)nslots;
    while (smem > (size_t)di.max_smem_optin && nslots > NW + 1) {
      --nslots;
      smem = 8 * (b_d + (size_t)nslots * slot_d) + 16 * (size_t)nslots;
    }
    if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
    for (int r = 0; r < nrows; ++r) {
      const double* u = static_cast<const double*>(rows.field[r]);
      double* out = static_cast<double*>(rows.out[r]);
      if (is_div) {
        if (int rc = set_smem(k_div_dmma<NW>, smem)) return rc;
        k_div_dmma<NW><<<(unsigned)grid, threads, smem, st>>>(J, O, u, out, E, nslots, tma_ok);
      } else {
        if (int rc = set_smem(k_grad_dmma<NW>, smem)) return rc;
        k_grad_dmma<NW><<<(unsigned)grid, threads, smem, st>>>(J, O, u, out, E, nslots, tma_ok);
      }
      if (int rc = post_launch()) return rc;
    }
    return FNSM_OK;
  }
  size_t smem = 8 * ((size_t)LiftLayout::B_DOUBLES + (size_t)nslots * LiftLayout::SLOT_DOUBLES) + 16 * (size_t)nslots;
  if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  if (kind == FNSM_OP_LIFT_FE) {
    if (int rc = set_smem(k_lift_dmma<NW, true>, smem)) return rc;
    k_lift_dmma<NW, true><<<(unsigned)grid, threads, smem, st>>>(J, O, rows, nrows, E, nslots, tma_ok);
  } else {
    if (int rc = set_smem(k_lift_dmma<NW, false>, smem)) return rc;
    k_lift_dmma<NW, false><<<(unsigned)grid, threads, smem, st>>>(J, O, rows, nrows, E, nslots, tma_ok);
  }
  return post_launch();
}

// wave_3d_p4: placeholder sequencing of the three dmma kernels until the
// single-launch kernel lands (tracked in DESIGN.md)
static int launch_wave3d_dmma(const fnsm_wave_args* a, long long E, const fnsm_cfg* cfg,
                              const DevInfo& di, cudaStream_t st) {
  OpmatRows r1{}; r1.field[0] = a->v; r1.out[0] = a->div_out;
  if (int rc = launch_dmma(FNSM_OP_DIV, a->J, a->D, r1, 1, 3, 35, 35, E, cfg, di, st)) return rc;
  OpmatRows r2{}; r2.field[0] = a->u; r2.out[0] = a->grad_out;
  if (int rc = launch_dmma(FNSM_OP_GRAD, a->J, a->D, r2, 1, 3, 35, 35, E, cfg, di, st)) return rc;
  OpmatRows r3{};
  for (int k = 0; k < 4; ++k) { r3.field[k] = a->F[k]; r3.out[k] = a->lift_out[k]; }
  return launch_dmma(FNSM_OP_LIFT_FE, a->Jface, a->L, r3, 4, 4, 35, 15, E, cfg, di, st);
}

}  // namespace fnsm
