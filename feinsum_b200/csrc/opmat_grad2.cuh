// DG gradient, fp64, p = 4 tets -- second formulation ("grad2"): the divergence kernel's operator tables, direct stores.
//
//   T[r][e][i] = sum_j D[r,i,j] u[e,j]          (three 35 x 35 applications to the same u)
//   out[x,e,i] = sum_r J[x,r,e] T[r][e][i]
//
// k_grad_dmma (opmat_dmma.cuh) lays the N = (dof, r) = 105 columns out in 14 column tiles so that a lane ends up with
// whole r-triples; that needs its own operator table (32 KB), 252 DMMAs per chunk (105 / 112 useful columns), and an
// output stage of 13 KB per warp because a lane's triples are scattered over the rows.  Here the column tiles are
// (r, nt): tile nt of operator r -- exactly the fragments of the DIVERGENCE table, sB[(kt = 3 jq + r, nt)] =
// D[r][8 nt + g][4 jq + t] -- so after the k-loop a lane holds T[0..2][e][8 nt + 2 t, +1] for its element: it applies J
// in registers and writes 64 contiguous bytes per element row straight to global memory, like the divergence does.
//   * 216 DMMAs + 162 DFMAs (dofs 32..34, from the divergence's left-over table) instead of 252 DMMAs: 6 % less
//     FP64-pipe time;
//   * no output stage: 5.6 KB of shared memory per warp instead of 19 KB -> 12 warps per SM instead of 10;
//   * the operator tables are shared with k_div_dmma: a fused div + grad kernel stages them once (opmat_wave.cuh).
#pragma once
#include "opmat_dmma.cuh"

namespace fnsm {

// stage the divergence operator tables (main fragments + left-over dofs) of D(3,35,35) into shared memory
__device__ __forceinline__ void stage_div_tables(double* sB, double* sL, const double* __restrict__ Dg) {
  using T = DivLayout;
  _Pragma("unroll 4")
  for (int idx = threadIdx.x; idx < T::B_MAIN; idx += blockDim.x) {
    const int h = idx & 1, ln = (idx >> 1) & 31, p = (idx >> 6) & 1, kt = idx >> 7;
    const int g = ln >> 2, t = ln & 3, jq = kt / 3, r = kt - 3 * jq;
    const int i = 8 * (2 * p + h) + g, j = 4 * jq + t;
    sB[idx] = (j < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.0;
  }
  _Pragma("unroll 4")
  for (int idx = threadIdx.x; idx < T::B_LEFT; idx += blockDim.x) {
    const int d = idx & 3, t = (idx >> 2) & 3, kt = idx >> 4;
    const int jq = kt / 3, r = kt - 3 * jq, j = 4 * jq + t;
    sL[idx] = (d < kNL && j < 35) ? Dg[(r * 35 + 32 + d) * 35 + j] : 0.0;
  }
}

// One gradient chunk from a slot that holds u[16][35] (at su) and J[9][16] (at sJ) -- shared by k_grad2_dmma and the
// fused wave kernel.  `after_load(step)` is called once behind every (jq, r) step of the two DMMA passes (54 steps):
// the plain producers weave their copies in there.
struct NoHook { __device__ __forceinline__ void operator()(int) const {} };

template <class Hook>
__device__ __forceinline__ void grad2_compute(const double (&a)[kME][9], const double (&Jr)[kME][9],
                                              uint32_t bB, uint32_t bL, double* __restrict__ outg, long long e0,
                                              long long E, int g, int t, Hook&& hook) {
  // ---- two passes over the column tiles (nt = 2 p, 2 p + 1), 108 DMMAs each ----
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    double acc[kME][3][2][2];
#pragma unroll
    for (int m = 0; m < kME; ++m)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h) { acc[m][r][h][0] = 0.0; acc[m][r][h][1] = 0.0; }
    double2 bq[2];
    bq[0] = lds_v2(bB + p * 512);
#pragma unroll
    for (int kt = 0; kt < 27; ++kt) {                     // kt = 3 jq + r
      const int jq = kt / 3, r = kt - 3 * jq, c = kt & 1;
      if (kt + 1 < 27) bq[c ^ 1] = lds_v2(bB + ((kt + 1) * 2 + p) * 512);
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        dmma884(acc[m][r][0], a[m][jq], bq[c].x);
        dmma884(acc[m][r][1], a[m][jq], bq[c].y);
      }
      hook(p * 27 + kt);
    }
    // J applied in registers, 16 bytes per (x, element, column tile) straight to global memory
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const long long e = e0 + chunk_el(g, m);
      if (e < E) {
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          double* o = outg + ((long long)x * E + e) * 35 + 16 * p + 2 * t;
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < 2; ++c)
              stg_stream(o + 8 * h + c, fma(Jr[m][3 * x + 2], acc[m][2][h][c],
                                            fma(Jr[m][3 * x + 1], acc[m][1][h][c], Jr[m][3 * x] * acc[m][0][h][c])));
        }
      }
    }
  }
  // ---- dofs 32..34 of the three operators: DFMA from the same A registers, lane <-> k = t (mod 4) partial sums ----
  double accL[kME][3][kNL];
#pragma unroll
  for (int m = 0; m < kME; ++m)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < kNL; ++d) accL[m][r][d] = 0.0;
#pragma unroll
  for (int kt = 0; kt < 27; ++kt) {
    const int jq = kt / 3, r = kt - 3 * jq;
    const double2 l01 = lds_v2(bL + kt * 128), l2x = lds_v2(bL + kt * 128 + 16);
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      accL[m][r][0] = fma(a[m][jq], l01.x, accL[m][r][0]);
      accL[m][r][1] = fma(a[m][jq], l01.y, accL[m][r][1]);
      accL[m][r][2] = fma(a[m][jq], l2x.x, accL[m][r][2]);
    }
  }
  // reduce over the quad and scatter: lane t ends with the total of dof 32 + t (t < 3).  Step 1 (xor 2): lanes
  // t = 0, 1 keep dofs (0, 1) and hand dof 2 over, lanes t = 2, 3 keep dof 2 (and a dummy); step 2 (xor 1) splits
  // the pair.  3 shuffles per (m, r) instead of the 6 of three full quad sums.
#pragma unroll
  for (int m = 0; m < kME; ++m) {
    double Tl[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const bool hi = t >= 2;
      // step 1: exchange with lane t ^ 2
      const double send_a = hi ? accL[m][r][0] : accL[m][r][2];
      const double send_b = accL[m][r][1];                         // only the hi lanes' copy is used by the receiver
      const double got_a = __shfl_xor_sync(0xffffffffu, send_a, 2);
      const double got_b = __shfl_xor_sync(0xffffffffu, send_b, 2);
      // lo lanes (t = 0, 1) now own dofs 0, 1; hi lanes (t = 2, 3) own dof 2
      const double k0 = hi ? accL[m][r][2] + got_a : accL[m][r][0] + got_a;   // lo: dof 0, hi: dof 2
      const double k1 = hi ? 0.0 : accL[m][r][1] + got_b;                     // lo: dof 1
      // step 2: exchange with lane t ^ 1: even lanes keep k0, odd lanes keep k1 (lo) / nothing (hi)
      const bool odd = t & 1;
      const double send2 = hi ? k0 : (odd ? k0 : k1);
      const double got2 = __shfl_xor_sync(0xffffffffu, send2, 1);
      Tl[r] = hi ? k0 + got2 : (odd ? k1 + got2 : k0 + got2);      // t = 0: dof 0, t = 1: dof 1, t = 2, 3: dof 2
    }
    const long long e = e0 + chunk_el(g, m);
    if (e < E && t < kNL) {
#pragma unroll
      for (int x = 0; x < 3; ++x)
        stg_stream(outg + ((long long)x * E + e) * 35 + 32 + t,
                   fma(Jr[m][3 * x + 2], Tl[2], fma(Jr[m][3 * x + 1], Tl[1], Jr[m][3 * x] * Tl[0])));
    }
  }
}

// plain (non-TMA) producer of this experiment: 8-byte cp.async from every lane, completing on the slot's barrier
// (32 arrivals).  k_grad_dmma has since moved to 1-D bulk copies (grad_issue_bulk); this kernel keeps the first form.
__device__ __forceinline__ void grad2_issue_plain(double* s, uint64_t* bar, const double* __restrict__ Jg,
                                                  const double* __restrict__ ug, long long e0, long long E, int lane) {
  using L = GradLayout;
  const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
  cp_async_run<L::U_SLAB>(s, ug + e0 * 35, ne * 35, lane);
  const int el = lane & (kCH - 1);
  const double* src = Jg + (long long)(lane >> 4) * E + e0 + el;
#pragma unroll
  for (int q = 0; q < (9 * kCH + 31) / 32; ++q)
    if (lane + 32 * q < 9 * kCH && el < ne) cp_async8(s + L::U_SLAB + lane + 32 * q, src + (long long)(2 * q) * E);
  cp_async_arrive_noinc(bar);
}
struct GradPlainCtx {
  uint32_t s;
  const double* u;            // next chunk + lane
  const double* j;            // row (lane >> 4), element e0 + (lane & 15)
  long long E2;
  int nvalid, ne, lane;
  bool active;
};

template <int NW, bool TMA = true>
__global__ void __launch_bounds__(NW * 32, 1)
k_grad2_dmma(const __grid_constant__ OpMaps maps, const double* __restrict__ Jg, const double* __restrict__ Dg,
             const double* __restrict__ ug, double* __restrict__ outg, long long E, int flags) {
  using L = GradLayout;                  // slot: u[16][35] + J[9][16]
  using T = DivLayout;                   // operator tables
  release_dependent_kernels();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* sL = sB + T::B_MAIN;
  double* slots = sB + T::B_DOUBLES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + (size_t)NW * L::SLOT_DOUBLES);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], TMA ? 1 : 32);
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();

  double* s = slots + (size_t)warp * L::SLOT_DOUBLES;
  uint64_t* bar = &bars[warp];
  const double* sJ = s + L::U_SLAB;
  const long long nchunks = (E + kCH - 1) / kCH;
  const WorkQueue wq{work_ctr, nchunks};
  const int g = lane >> 2, t = lane & 3, tpad = t < 3 ? t : 2;
  const bool tma = TMA && (flags & kFlagTma);

  long long cur = wq.take(lane), nxt = wq.take(lane);
  if (cur < nchunks) {
    if constexpr (TMA) grad_issue(s, bar, &maps, Jg, ug, cur, E, tma, lane);
    else               grad2_issue_plain(s, bar, Jg, ug, cur * kCH, E, lane);
  }
  stage_div_tables(sB, sL, Dg);
  __syncthreads();
  const uint32_t bB = smem_u32(sB) + lane * 16, bL = smem_u32(sL) + t * 32;
  for (uint32_t n = 0; cur < nchunks; ++n) {
    mbar_wait(bar, n & 1u);
    double a[kME][9], Jr[kME][9];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
#pragma unroll
      for (int jq = 0; jq < 9; ++jq)           // k-slot j = 35 is padding (zero operator entries): reads j = 34
        a[m][jq] = s[el * 35 + (jq == 8 ? 32 + tpad : 4 * jq + t)];
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jr[m][xr] = sJ[xr * kCH + el];
    }
    __syncwarp();
    const long long e0 = cur * kCH;
    if constexpr (TMA) {
      if (nxt < nchunks) grad_issue(s, bar, &maps, Jg, ug, nxt, E, tma, lane);
      const unsigned tk = wq.ticket(lane);
      grad2_compute(a, Jr, bB, bL, outg, e0, E, g, t, NoHook{});
      cur = nxt;
      nxt = wq.resolve(tk);
    } else {
      GradPlainCtx pc;
      const long long e0n = nxt * kCH;
      pc.active = nxt < nchunks;
      pc.ne = (int)((E - e0n < kCH) ? (E - e0n) : kCH);
      pc.nvalid = pc.ne * 35;
      pc.lane = lane;
      pc.s = smem_u32(s + lane);
      pc.E2 = 2 * E;
      pc.u = ug + e0n * 35 + lane;
      pc.j = Jg + (long long)(lane >> 4) * E + e0n + (lane & (kCH - 1));
      const unsigned tk = wq.ticket(lane);
      // 18 rows of u on steps 0, 2, .., 34, the 5 rows of J on steps 36, 38, .., 44 (steps are compile-time after unrolling)
      grad2_compute(a, Jr, bB, bL, outg, e0, E, g, t, [&](int step) {
        if (!pc.active || (step & 1)) return;
        const int q = step >> 1;
        if (q < 18) {
          if (pc.lane + 32 * q < pc.nvalid) cp_async8_rr(pc.s + 32 * q * 8, pc.u + 32 * q);
        } else if (q < 23) {
          const int r = q - 18;
          if (pc.lane + 32 * r < 9 * kCH && (pc.lane & (kCH - 1)) < pc.ne)
            cp_async8_rr(pc.s + (L::U_SLAB + 32 * r) * 8, pc.j + (long long)r * pc.E2);
        }
      });
      if (pc.active) cp_async_arrive_noinc(bar);
      cur = nxt;
      nxt = wq.resolve(tk);
    }
  }
  wait_for_previous_kernels();
}

template <int NW>
static int launch_grad2(const double* J, const double* D, const OpmatRows& rows, int nrows, long long E,
                        bool force_plain, const DevInfo& di, cudaStream_t st) {
  using L = GradLayout;
  using T = DivLayout;
  const size_t smem = 8 * ((size_t)T::B_DOUBLES + (size_t)NW * L::SLOT_DOUBLES) + 8 * (size_t)NW + 8;
  if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  const long long nchunks = (E + kCH - 1) / kCH;
  const long long need = (nchunks + NW - 1) / NW;
  const unsigned grid = (unsigned)(di.sms < need ? di.sms : need);
  if (int rc = set_smem(k_grad2_dmma<NW, true>, smem)) return rc;
  if (int rc = set_smem(k_grad2_dmma<NW, false>, smem)) return rc;
  for (int r = 0; r < nrows; ++r) {
    const double* u = static_cast<const double*>(rows.field[r]);
    double* out = static_cast<double*>(rows.out[r]);
    OpMaps maps{};
    const bool tma = !force_plain && E % 2 == 0 && E < (1LL << 31) - kCH && aligned16(J) && aligned16(u) &&
                     map_erows(&maps.jac, J, E, 9) && map_rows(&maps.in, u, E, 35);
    if (tma) launch_k(k_grad2_dmma<NW, true>, grid, NW * 32, smem, st, maps, J, D, u, out, E, (int)kFlagTma);
    else     launch_k(k_grad2_dmma<NW, false>, grid, NW * 32, smem, st, maps, J, D, u, out, E, 0);
    if (int rc = post_launch()) return rc;
  }
  return FNSM_OK;
}

}  // namespace fnsm
