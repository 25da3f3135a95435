// fp32 DG operator kernels on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators
// in TMEM), 3xTF32: both operands are split hi = tf32(x), lo = tf32(x - hi) and
//     main  = A_hi B_hi            (its own fp32 accumulator)
//     corr  = A_hi B_lo + A_lo B_hi (a second accumulator; the lo*lo term is dropped)
// restores ~2^-21 relative accuracy per product (opmat_tf32.cuh, the mma.sync variant, explains why
// the cross terms get their own accumulator).  The legacy mma.sync path tops out at 277 TFLOP/s TF32
// (tools/ubench4) -> its 3xTF32 kernels are capped at ~89 % of the fp32 roofline by the tensor pipe
// alone; tcgen05 lifts that cap, the kernels below are bounded by HBM.
//
// Shape of one tile: M = 128 elements (TMEM lane = element), K = contracted dofs, N = output columns.
//   grad:  C[e][(i,r)] = sum_j u[e,j] D[r,i,j]   K = 35 -> 40, N = 105 -> 112, column n = 3 i + r;
//          out[x,e,i] = sum_r J[x,r,e] C[e][(i,r)] is applied by the thread that owns element e
//          (tcgen05.ld 32x32b: one TMEM lane per thread) -- no shuffles, no shared-memory transpose.
//   One MMA per k-step computes  A_hi x [B_hi | pad | B_lo]  (N = 240: columns 0..111 main,
//   128..239 corr), a second one (N = 112) adds  A_lo x B_hi  into the corr columns.
//
// Structure: one persistent CTA per SM with several independent GROUPS of warps -- two at p = 4, three to
// eight at the lower orders, whose tiles are small enough (6-40 KB) that the fixed latencies of a tile
// dominate.  A group owns its TMA slot(s), its share of the 512 TMEM columns, an output stage (grad: the
// A_hi / A_lo operand buffer aliases it) and its mbarriers, and walks its tiles serially:
//     wait slot -> split rows (grad: into the UMMA canonical K-major layout in shared memory; div / lift:
//     straight into TMEM) -> elected thread issues the MMAs + tcgen05.commit and re-arms the slot with
//     the next TMA load -> wait commit -> tcgen05.ld, apply J (grad), stage -> one TMA tensor store.
// While one group waits on its loads / MMAs / stores the others convert or drain, so the SM stays busy
// without any cross-group synchronisation.  A group is 4 warps (warp w owns TMEM lanes 32 (w & 3) .. +31,
// thread = element row) or, for grad at p = 4, 8 warps: warps w and w + 4 share a lane quadrant and split
// the k-columns of the conversion and the dof columns of the epilogue (that pays where conversion and
// epilogue are long -- grad 90 -> 99 % -- and costs where a tile already has several barriers -- div
// 94 -> 90 %, lift 95 -> 93 %).
#pragma once
#include "opmat_tf32.cuh"

namespace fnsm {

// ------------------------------------------------------------- tcgen05 -----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B stored as 128
// contiguous bytes; LBO = byte distance between the two core matrices of one k-step (K direction),
// SBO = byte distance between consecutive 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A lives in TMEM (lane = row, one fp32 column per k); issued by ONE thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
// 8 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------- plain (non-TMA) producer / drain -----
// Operands that do not qualify for tensor maps (E % 4 != 0: the rows of every array whose element axis is not
// leading start at addresses that are not multiples of 16; or a base pointer that is not 16-byte aligned) used to fall
// off the tcgen05 kernels onto the mma.sync ones (63-77 % of roofline instead of 94-99 %).  They now run the SAME
// kernels with TMA = false: the threads of a group move a tile's rows with cp.async -- still asynchronous, still
// completing on the slot's mbarrier (cp.async.mbarrier.arrive) -- and drain the stage with plain vector stores.
// A slab lands at  slab_base + (global address & 15):  source and destination are congruent mod 16, so everything
// but the ragged ends of a slab moves in 16-byte pieces; consumers read the slot with scalar 4-byte accesses
// (rows of 35 / 15 floats), so the 0 / 4 / 8 / 12-byte shift costs them nothing.  TMA clipping / zero fill of the
// tail tile becomes "copy only the valid rows" (stale rows are computed -- tile rows are independent -- and not stored).
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// one arrival (counted in the barrier's expected count) once all earlier cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int shift16(const void* p) { return (int)(reinterpret_cast<uintptr_t>(p) & 15); }
// tx bytes added to the pending phase without an arrival (the arrivals are the threads' cp_async_arrive)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// nbytes (multiple of 4) from the 4-byte aligned global address g to s_base + shift16(g); s_base is 16-byte aligned.
// `whole` (every tile but the last of the arrays): ONE 1-D bulk copy by thread 0 of the group, from g rounded down to
// g + nbytes rounded up to 16 bytes -- the few bytes in front and behind belong to the neighbouring tiles / slabs --
// with the bytes announced on `bar`, where every thread of the group still arrives (cp_async_arrive).  The cooperative
// cp.async form below (~9 LDGSTS per thread and slab behind a rolled, branchy loop) is what the last tile takes: it
// is partial, and rounding its end up would read past the arrays.
template <int NT>
__device__ __forceinline__ void plain_load_async(unsigned char* s_base, const void* gp, int nbytes, int tid,
                                                 uint64_t* bar, bool whole) {
  const unsigned char* g = static_cast<const unsigned char*>(gp);
  const int a = shift16(g);
  if (whole) {
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(a + nbytes + 15) & ~15u;
      mbar_expect_tx(bar, bytes);
      bulk_load(s_base, g - a, bytes, bar);
    }
    return;
  }
  const uint32_t s = smem_u32(s_base) + a;
  int head = (16 - a) & 15;
  if (head > nbytes) head = nbytes;
  if (tid < (head >> 2)) cp_async4(s + 4 * tid, g + 4 * tid);
  const int body = (nbytes - head) & ~15, end = head + body;
  for (int o = head + 16 * tid; o < end; o += 16 * NT) cp_async16(s + o, g + o);
  if (tid < ((nbytes - end) >> 2)) cp_async4(s + end + 4 * tid, g + end + 4 * tid);
}
// the reverse: nbytes staged at s_base + shift16(g) -> global g: the 16-byte aligned middle as one bulk store by
// thread 0 (bulk async-group: tma_store_wait_read before the stage is rewritten), the ragged ends (< 16 bytes each)
// by the first threads.  The caller has fenced (fence.proxy.async) and synchronised the group.
template <int NT>
__device__ __forceinline__ void plain_store(void* gp, const unsigned char* s_base, int nbytes, int tid) {
  unsigned char* g = static_cast<unsigned char*>(gp);
  const int a = shift16(g);
  const unsigned char* s = s_base + a;
  int head = (16 - a) & 15;
  if (head > nbytes) head = nbytes;
  const int body = (nbytes - head) & ~15, end = head + body;
  if (tid == 0) {
    if (body > 0) bulk_store(g + head, s + head, (uint32_t)body);
    tma_store_commit();
  }
  if (tid >= 32 && tid < 32 + (head >> 2)) {
    const int k = tid - 32;
    *reinterpret_cast<float*>(g + 4 * k) = *reinterpret_cast<const float*>(s + 4 * k);
  }
  if (tid >= 64 && tid < 64 + ((nbytes - end) >> 2)) {
    const int k = tid - 64;
    *reinterpret_cast<float*>(g + end + 4 * k) = *reinterpret_cast<const float*>(s + end + 4 * k);
  }
}
constexpr int kPlainPad = 128;      // slack per slab of a slot / stage in the non-TMA kernels (shift <= 12 B; keeps 128-B alignment)

// ================================================================ GRAD =====

constexpr int tc_pad(int v, int m) { return (v + m - 1) / m * m; }
constexpr int tc_max(int a, int b) { return a > b ? a : b; }

// ND = volume dofs per element (p = 1..4 tets: 4, 10, 20, 35); sizes in the comments are for ND = 35
template <int ND>
struct GradTC {
  static constexpr int TM = 128;                        // elements per tile
  static constexpr int K = tc_pad(ND, 8), KS = K / 8;   // padded contraction length (40), k-steps of 8
  static constexpr int N = tc_pad(3 * ND, 16);          // columns n = 3 i + r (112, 105 used)
  static constexpr int NP = tc_pad(N, 32);              // column pitch main -> corr (TMEM) = row pitch hi -> lo (table)
  static constexpr int NB = NP + N;                     // rows of the operator table: [hi | pad | lo] = 240
  // groups per CTA: the tiles of the lower orders are small (6-40 KB), so their fixed latencies (TMA, MMA
  // completion, group barriers) are covered by four independent groups instead of two
  // p = 2 (ND = 10): 4 / 6 / 7 / 8 groups = 79 / 90.5 / 88.8 / 86 % of roofline (8 groups = 1024 threads cap the
  // kernel at 64 registers: 40 bytes of spills)
  static constexpr int GROUPS = ND <= 4 ? 8 : (ND <= 10 ? 6 : (ND <= 20 ? 4 : 2));
  static constexpr int WPG = GROUPS == 2 ? 8 : 4, GT = 32 * WPG, NH = WPG / 4;   // warps / threads per group, threads per row
  static constexpr int THREADS = GROUPS * GT;
  static constexpr int B_LBO = NB * 16;                 // operator table: addr(n, k) = (k/4) B_LBO + 16 n + 4 (k%4)
  static constexpr int A_LBO = TM * 16;                 // A operand:      addr(e, k) = (k/4) A_LBO + 16 e + 4 (k%4)
  static constexpr int B_BYTES = (K / 4) * B_LBO;       // 35 840
  static constexpr int A_BYTES = (K / 4) * A_LBO;       // 20 480 per half (hi, lo)
  static constexpr int SLOT_BYTES = TM * ND * 4;        // 17 920: u rows of one tile
  static constexpr int OUT_BYTES = 3 * TM * ND * 4;     // 53 760: out[x][e][i]
  static constexpr int STAGE_BYTES = tc_pad(tc_max(OUT_BYTES, 2 * A_BYTES), 128);   // the A operand aliases the stage
  static constexpr int NQ = (ND + 7) / 8;               // epilogue passes of 8 dofs (24 columns)
  static constexpr int GROUP_BYTES = 2 * SLOT_BYTES + STAGE_BYTES;
  static constexpr int TMEM_COLS_PER_GROUP = (512 / GROUPS) / 16 * 16;      // 256 (240 used)
  static constexpr size_t SMEM = B_BYTES + (size_t)GROUPS * GROUP_BYTES + 512;   // + mbarriers, TMEM base
  // non-TMA variant: every slab of the slots and of the stage carries kPlainPad bytes of slack for its shift
  static constexpr int OUT_SLAB_P = TM * ND * 4 + kPlainPad;
  static constexpr int SLOT_P = SLOT_BYTES + kPlainPad;
  static constexpr int STAGE_P = tc_pad(tc_max(3 * OUT_SLAB_P, 2 * A_BYTES), 128);
  static constexpr int GROUP_P = 2 * SLOT_P + STAGE_P;
  static constexpr size_t SMEM_P = B_BYTES + (size_t)GROUPS * GROUP_P + 512;
  static_assert(NB <= TMEM_COLS_PER_GROUP && NP + tc_pad(3 * ND, 8) <= TMEM_COLS_PER_GROUP, "TMEM budget");
  static_assert(SMEM_P <= 227 * 1024, "shared-memory budget");
  static_assert(SLOT_BYTES % 128 == 0, "TMA destination alignment");
};

struct GradTCMaps { CUtensorMap in, out; };

// TMA = false: plain producer / drain (see "plain (non-TMA) producer" above); ug / outg are only used there
template <int ND, bool TMA = true>
__global__ void __launch_bounds__(GradTC<ND>::THREADS, 1)
k_grad_tc32(const __grid_constant__ GradTCMaps maps, const float* __restrict__ Jg, const float* __restrict__ Dg,
            const float* __restrict__ ug, float* __restrict__ outg, long long E) {
  using L = GradTC<ND>;
  constexpr int GROUP_BYTES = TMA ? L::GROUP_BYTES : L::GROUP_P, SLOT_PITCH = TMA ? L::SLOT_BYTES : L::SLOT_P;
  constexpr int OUT_SLAB = TMA ? L::TM * ND * 4 : L::OUT_SLAB_P;     // bytes between the x-slabs of the stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sB = smem_raw;
  unsigned char* groups = smem_raw + L::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(groups + (size_t)L::GROUPS * GROUP_BYTES);   // [group][full0, full1, mma]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * L::GROUPS);

  const int warp = uniform_warp_idx();
  const int gq = warp / L::WPG, wq = warp - gq * L::WPG; // group, warp in group
  const int row = (wq & 3) * 32 + (threadIdx.x & 31);    // element row of the tile = TMEM lane
  const int half = L::NH == 1 ? 0 : wq >> 2;             // which share of the row's columns (constant 0 for WPG = 4)
  const bool leader = wq == 0 && (threadIdx.x & 31) == 0;
  const int gtid = (int)threadIdx.x - gq * L::GT;        // thread index inside the group

  if (threadIdx.x == 0) {
    // plain producer: every thread of the group arrives on a slot's barrier (through its cp.async group)
    for (int k = 0; k < 3 * L::GROUPS; ++k) mbar_init(&bars[k], (!TMA && k % 3 != 2) ? L::GT : 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // operator table [hi | pad | lo], UMMA K-major canonical layout, zero padded (k >= 35, n >= 105)
  _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
  for (int idx = threadIdx.x; idx < L::NP * L::K; idx += blockDim.x) {
    const int n = idx / L::K, k = idx - n * L::K;
    const int i = n / 3, r = n - 3 * i;
    const float v = (n < 3 * ND && k < ND) ? Dg[(r * ND + i) * ND + k] : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int off = (k >> 2) * L::B_LBO + (k & 3) * 4;
    *reinterpret_cast<uint32_t*>(sB + off + n * 16) = hi;
    if (n < L::N) *reinterpret_cast<uint32_t*>(sB + off + (L::NP + n) * 16) = lo;
  }
  fence_proxy_async();                                   // table is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot + (uint32_t)gq * L::TMEM_COLS_PER_GROUP;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp & 3) * 32u << 16);

  unsigned char* gbase = groups + (size_t)gq * GROUP_BYTES;
  unsigned char* slot_b[2] = {gbase, gbase + SLOT_PITCH};
  const int ushift = TMA ? 0 : shift16(ug);              // tile rows start at multiples of 512 ND bytes: one shift for all tiles
  const float* slot[2] = {reinterpret_cast<const float*>(slot_b[0] + ushift), reinterpret_cast<const float*>(slot_b[1] + ushift)};
  unsigned char* stage_b = gbase + 2 * SLOT_PITCH;
  float* stage_x[3];
#pragma unroll
  for (int x = 0; x < 3; ++x)
    stage_x[x] = reinterpret_cast<float*>(stage_b + x * OUT_SLAB + (TMA ? 0 : shift16(outg + (long long)x * E * ND)));
  uint64_t* full = &bars[3 * gq];
  uint64_t* mma_done = &bars[3 * gq + 2];
  const int bar_id = 1 + gq;

  const long long ntiles = (E + L::TM - 1) / L::TM;
  const long long tile0 = (long long)blockIdx.x * L::GROUPS + gq, tstride = (long long)gridDim.x * L::GROUPS;
  constexpr uint32_t idesc_wide = umma_idesc_tf32(L::TM, L::NB), idesc_half = umma_idesc_tf32(L::TM, L::N);

  // fetch tile tl into slot p.  TMA: the leader; plain: every thread of the group moves its share
  auto fetch = [&](long long tl, int p) {
    if constexpr (TMA) {
      if (leader) {
        mbar_arrive_expect_tx(&full[p], L::SLOT_BYTES);
        tma_load_2d(slot_b[p], &maps.in, 0, (int)(tl * (L::TM / 4)), &full[p]);
      }
    } else {
      const long long e0 = tl * L::TM;
      const int rows = (int)(E - e0 < L::TM ? E - e0 : L::TM);
      plain_load_async<L::GT>(slot_b[p], ug + e0 * ND, rows * ND * 4, gtid, &full[p], tl + 1 < ntiles);
      cp_async_arrive(&full[p]);
    }
  };
#pragma unroll
  for (int p = 0; p < 2; ++p)
    if (tile0 + p * tstride < ntiles) fetch(tile0 + p * tstride, p);
  // Jacobian of this thread's element, fetched one tile ahead (coalesced: J[xr][e])
  float Jn[9];
  {
    const long long e = tile0 * L::TM + row;
#pragma unroll
    for (int xr = 0; xr < 9; ++xr) Jn[xr] = (tile0 < ntiles && e < E) ? __ldg(Jg + (long long)xr * E + e) : 0.f;
  }

  uint32_t it = 0;
  for (long long tile = tile0; tile < ntiles; tile += tstride, ++it) {
    const int s = it & 1;
    float Jr[9];
#pragma unroll
    for (int xr = 0; xr < 9; ++xr) Jr[xr] = Jn[xr];
    {
      const long long tn = tile + tstride, e = tn * L::TM + row;
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jn[xr] = (tn < ntiles && e < E) ? __ldg(Jg + (long long)xr * E + e) : 0.f;
    }
    mbar_wait(&full[s], (it >> 1) & 1u);
    // the A operand aliases the output stage: the previous tile's bulk store must have read it out
    // (bulk drain: the leader waits for its bulk store of the previous tile to have read the stage out)
    if (leader) tma_store_wait_read();
    group_barrier(bar_id, L::GT);
    // ---- element row -> A_hi / A_lo (K-major canonical layout: 16-byte k-quads, rows 16 B apart) ----
    {
      const float* su = slot[s] + row * ND;
      unsigned char* ahi = stage_b + row * 16;
      unsigned char* alo = ahi + L::A_BYTES;
#pragma unroll
      for (int kq = 0; kq < L::K / 4; ++kq) {
        if (L::NH > 1 && (kq % L::NH) != half) continue;   // warp-uniform; kq stays a compile-time constant
        uint4 h, l;
        split_tf32(4 * kq + 0 < ND ? su[4 * kq + 0] : 0.f, h.x, l.x);
        split_tf32(4 * kq + 1 < ND ? su[4 * kq + 1] : 0.f, h.y, l.y);
        split_tf32(4 * kq + 2 < ND ? su[4 * kq + 2] : 0.f, h.z, l.z);
        split_tf32(4 * kq + 3 < ND ? su[4 * kq + 3] : 0.f, h.w, l.w);
        *reinterpret_cast<uint4*>(ahi + kq * L::A_LBO) = h;
        *reinterpret_cast<uint4*>(alo + kq * L::A_LBO) = l;
      }
    }
    fence_proxy_async();                                 // generic-proxy writes -> tensor core / TMA reads
    tc_fence_before();
    group_barrier(bar_id, L::GT);
    if (leader) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(stage_b), a_lo = a_hi + L::A_BYTES, b0 = smem_u32(sB);
#pragma unroll
      for (int ks = 0; ks < L::KS; ++ks) {
        const uint64_t dah = umma_desc(a_hi + ks * 2 * L::A_LBO, L::A_LBO, 128);
        const uint64_t dal = umma_desc(a_lo + ks * 2 * L::A_LBO, L::A_LBO, 128);
        const uint64_t db = umma_desc(b0 + ks * 2 * L::B_LBO, L::B_LBO, 128);
        umma_tf32(tmem_base, dah, db, idesc_wide, ks > 0);            // main | corr  (+)= A_hi [B_hi | B_lo]
        umma_tf32(tmem_base + L::NP, dal, db, idesc_half, true);      // corr += A_lo B_hi
      }
      umma_commit(mma_done);
    }
    // the slot has been consumed by every thread of the group: fetch the tile two steps ahead
    if (tile + 2 * tstride < ntiles) fetch(tile + 2 * tstride, s);
    mbar_wait(mma_done, it & 1u);
    tc_fence_after();
    // ---- TMEM -> registers (8 dofs = 24 columns at a time), J applied, staged as out[x][e][i] ----
#pragma unroll
    for (int q = 0; q < L::NQ; ++q) {
      if (L::NH > 1 && (q % L::NH) != half) continue;   // warp-uniform; q stays a compile-time constant
      float m[24], c[24];
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        if (24 * q + 8 * p >= 3 * ND) continue;          // nothing useful (and possibly nothing allocated) there
        tmem_ld8(tmem_lane + 24 * q + 8 * p, reinterpret_cast<float(&)[8]>(m[8 * p]));
        tmem_ld8(tmem_lane + L::NP + 24 * q + 8 * p, reinterpret_cast<float(&)[8]>(c[8 * p]));
      }
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const int i = 8 * q + d;
        if (i < ND) {
          const float T0 = m[3 * d] + c[3 * d], T1 = m[3 * d + 1] + c[3 * d + 1], T2 = m[3 * d + 2] + c[3 * d + 2];
#pragma unroll
          for (int x = 0; x < 3; ++x)
            stage_x[x][row * ND + i] = fmaf(Jr[3 * x + 2], T2, fmaf(Jr[3 * x + 1], T1, Jr[3 * x] * T0));
        }
      }
    }
    tc_fence_before();                                   // TMEM reads done before the next tile's MMAs
    fence_proxy_async();
    group_barrier(bar_id, L::GT);
    if constexpr (TMA) {
      if (leader) {
        tma_store_3d(&maps.out, stage_b, 0, (int)(tile * (L::TM / 4)), 0);
        tma_store_commit();
      }
    } else {
      const long long e0 = tile * L::TM;
      const int rows = (int)(E - e0 < L::TM ? E - e0 : L::TM);
#pragma unroll
      for (int x = 0; x < 3; ++x)
        plain_store<L::GT>(outg + ((long long)x * E + e0) * ND, stage_b + x * OUT_SLAB, rows * ND * 4, gtid);
    }
  }
  if (leader) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ================================================================ LIFT =====
// out_k[e,i] = sum_{f,j} Op(f,i,j) Jf(e,f) v_k[f,e,j]:  K = (f, j) = 60 -> 64, N = 35 -> 48.
// The A operand (Jf v, split hi / lo) is written straight into TMEM by the thread that owns the row
// (tcgen05.st) and consumed from there (tcgen05.mma with A in TMEM): no shared-memory round trip.
// TMEM columns of a group (p = 4): [0,48) main, [64,112) corr, [112,176) A_hi, [176,240) A_lo.
// Work item = (tile, field); the face Jacobian of a tile is fetched once for its fields.
// ND volume dofs, NFD dofs per face (p = 1..4 tets: 4/3, 10/6, 20/10, 35/15), 4 faces
template <int ND, int NFD>
struct LiftTC {
  static constexpr int TM = 128;
  static constexpr int K = tc_pad(4 * NFD, 8), KS = K / 8;                      // 64, 8
  static constexpr int N = tc_pad(ND, 16), NP = tc_pad(N, 32), NB = NP + N;     // 48, 64; table rows [hi | pad | lo] = 112
  static constexpr int WPG = 4, GT = 32 * WPG, NH = WPG / 4;
  static constexpr int GROUPS = ND <= 4 ? 6 : (ND <= 10 ? 4 : (ND <= 20 ? 3 : 2)), THREADS = GROUPS * GT;   // see GradTC
  static constexpr int B_LBO = NB * 16;
  static constexpr int B_BYTES = (K / 4) * B_LBO;       // 28 672
  static constexpr int V_SLAB = TM * NFD;               // floats per face
  static constexpr int SLOT_BYTES = 4 * V_SLAB * 4;     // 30 720
  static constexpr int STAGE_BYTES = TM * ND * 4;       // 17 920
  static constexpr int NQ = (ND + 7) / 8;
  static constexpr int GROUP_BYTES = 2 * SLOT_BYTES + STAGE_BYTES;
  static constexpr int TMEM_COLS_PER_GROUP = GROUPS == 2 ? 256 : (GROUPS == 3 ? 160 : (GROUPS == 4 ? 128 : 80));
  static constexpr int A_HI_COL = tc_pad(NB, 16), A_LO_COL = A_HI_COL + K;      // 112 -> [0,112) D, [112,176) A_hi, [176,240) A_lo
  static_assert(A_LO_COL + K <= TMEM_COLS_PER_GROUP && K % 8 == 0, "TMEM budget");
  static_assert(SLOT_BYTES % 128 == 0 && STAGE_BYTES % 128 == 0, "TMA alignment");
  static constexpr size_t SMEM = B_BYTES + (size_t)GROUPS * GROUP_BYTES + 512;   // + mbarriers, TMEM base
  // non-TMA variant (see GradTC): face slabs and the stage carry kPlainPad bytes of slack
  static constexpr int V_SLAB_P = V_SLAB * 4 + kPlainPad;                       // bytes per face slab
  static constexpr int SLOT_P = 4 * V_SLAB_P, STAGE_P = STAGE_BYTES + kPlainPad;
  static constexpr int GROUP_P = 2 * SLOT_P + STAGE_P;
  static constexpr size_t SMEM_P = B_BYTES + (size_t)GROUPS * GROUP_P + 512;
  static_assert(SMEM_P <= 227 * 1024, "shared-memory budget");
};

struct LiftTCMaps { CUtensorMap in[8]; CUtensorMap out[8]; };

template <int ND, int NFD, bool FE, bool TMA = true>
__global__ void __launch_bounds__(LiftTC<ND, NFD>::THREADS, 1)
k_lift_tc32(const __grid_constant__ LiftTCMaps maps, const float* __restrict__ Jg, const float* __restrict__ Og,
            const __grid_constant__ OpmatRows rows, int nrows, long long E) {
  using L = LiftTC<ND, NFD>;
  constexpr int GROUP_BYTES = TMA ? L::GROUP_BYTES : L::GROUP_P, SLOT_PITCH = TMA ? L::SLOT_BYTES : L::SLOT_P;
  constexpr int V_SLAB_B = TMA ? L::V_SLAB * 4 : L::V_SLAB_P;       // bytes between the face slabs of a slot
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sB = smem_raw;
  unsigned char* groups = smem_raw + L::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(groups + (size_t)L::GROUPS * GROUP_BYTES);   // [group][full0, full1, mma]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * L::GROUPS);

  const int warp = uniform_warp_idx();
  const int gq = warp / L::WPG, wq = warp - gq * L::WPG;
  const int row = (wq & 3) * 32 + (threadIdx.x & 31);
  const int half = L::NH == 1 ? 0 : wq >> 2;
  const bool leader = wq == 0 && (threadIdx.x & 31) == 0;
  const int gtid = (int)threadIdx.x - gq * L::GT;

  if (threadIdx.x == 0) {
    for (int k = 0; k < 3 * L::GROUPS; ++k) mbar_init(&bars[k], (!TMA && k % 3 != 2) ? L::GT : 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // operator table [hi | pad | lo]: row n = dof i, k = 15 f + j
  _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
  for (int idx = threadIdx.x; idx < L::NP * L::K; idx += blockDim.x) {
    const int n = idx / L::K, k = idx - n * L::K;
    const int f = k / NFD, j = k - NFD * f;
    float v = 0.f;
    if (n < ND && k < 4 * NFD) v = FE ? Og[(n * 4 + f) * NFD + j] : Og[(f * ND + n) * NFD + j];
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int off = (k >> 2) * L::B_LBO + (k & 3) * 4;
    *reinterpret_cast<uint32_t*>(sB + off + n * 16) = hi;
    if (n < L::N) *reinterpret_cast<uint32_t*>(sB + off + (L::NP + n) * 16) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot + (uint32_t)gq * L::TMEM_COLS_PER_GROUP;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp & 3) * 32u << 16);

  unsigned char* gbase = groups + (size_t)gq * GROUP_BYTES;
  unsigned char* slot_b[2] = {gbase, gbase + SLOT_PITCH};
  unsigned char* stage_b = gbase + 2 * SLOT_PITCH;
  uint64_t* full = &bars[3 * gq];
  uint64_t* mma_done = &bars[3 * gq + 2];
  const int bar_id = 1 + gq;

  const long long ntiles = (E + L::TM - 1) / L::TM;
  const long long tile0 = (long long)blockIdx.x * L::GROUPS + gq, tstride = (long long)gridDim.x * L::GROUPS;
  constexpr uint32_t idesc_wide = umma_idesc_tf32(L::TM, L::NB), idesc_half = umma_idesc_tf32(L::TM, L::N);
  // items of this group: (tile0, 0), (tile0, 1), ..., (tile0 + tstride, 0), ...; item index `it`
  const long long my_tiles = tile0 < ntiles ? (ntiles - tile0 + tstride - 1) / tstride : 0;
  const long long nitems = my_tiles * nrows;

  // TMA: the leader; plain: every thread of the group moves its share of the four face slabs
  // items are issued in order, so the (tile, field) of the next one to issue advance by counting (a 64-bit division
  // per item and thread is a visible share of a p = 1 tile)
  long long issue_tl = tile0;
  int issue_fld = 0;
  auto issue = [&](int s) {
    const long long tl = issue_tl;
    const int fld = issue_fld;
    if (++issue_fld == nrows) { issue_fld = 0; issue_tl += tstride; }
    if constexpr (TMA) {
      if (leader) {
        mbar_arrive_expect_tx(&full[s], L::SLOT_BYTES);
        tma_load_3d(slot_b[s], &maps.in[fld], 0, (int)(tl * (L::TM / 4)), 0, &full[s]);
      }
    } else {
      const long long e0 = tl * L::TM;
      const int nr = (int)(E - e0 < L::TM ? E - e0 : L::TM);
      const float* vg = static_cast<const float*>(rows.field[fld]);
#pragma unroll
      for (int f = 0; f < 4; ++f)
        plain_load_async<L::GT>(slot_b[s] + f * V_SLAB_B, vg + ((long long)f * E + e0) * NFD, nr * NFD * 4, gtid,
                                &full[s], tl + 1 < ntiles);
      cp_async_arrive(&full[s]);
    }
  };
  if (nitems > 0) issue(0);
  if (nitems > 1) issue(1);
  const bool j_vec = TMA || shift16(Jg) == 0;            // J(E,4) rows are 16 bytes: vector load iff the base is aligned
  auto load_j = [&](long long tl, float (&J)[4]) {
    const long long e = tl * L::TM + row;
    if (tl < ntiles && e < E) {
      if (FE) {
#pragma unroll
        for (int f = 0; f < 4; ++f) J[f] = __ldg(Jg + (long long)f * E + e);
      } else if (j_vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(Jg) + e);
        J[0] = v.x; J[1] = v.y; J[2] = v.z; J[3] = v.w;
      } else {
#pragma unroll
        for (int f = 0; f < 4; ++f) J[f] = __ldg(Jg + e * 4 + f);
      }
    } else {
#pragma unroll
      for (int f = 0; f < 4; ++f) J[f] = 0.f;
    }
  };
  float Jf[4], Jn[4];
  load_j(tile0, Jn);

  int fld = 0;
  long long tile = tile0;
  for (long long it = 0; it < nitems; ++it) {
    const int s = (int)(it & 1);
    if (fld == 0) {
#pragma unroll
      for (int f = 0; f < 4; ++f) Jf[f] = Jn[f];
    }
    mbar_wait(&full[s], (uint32_t)(it >> 1) & 1u);
    // ---- row of the slot -> A_hi / A_lo in TMEM ----
    {
      const float* svf[4];                     // this row in the four face slabs (plain: each slab has its own shift)
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int sh = TMA ? 0 : shift16(static_cast<const float*>(rows.field[fld]) + (long long)f * E * NFD);
        svf[f] = reinterpret_cast<const float*>(slot_b[s] + f * V_SLAB_B + sh) + row * NFD;
      }
#pragma unroll
      for (int c = 0; c < L::K / 8; ++c) {     // 8 k at a time
        if (L::NH > 1 && (c % L::NH) != half) continue;   // warp-uniform; c stays a compile-time constant
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = 8 * c + q, f = k / NFD, j = k - NFD * f;
          const float a = k < 4 * NFD ? Jf[f < 4 ? f : 3] * svf[f < 4 ? f : 3][j] : 0.f;
          split_tf32(a, hi[q], lo[q]);
        }
        tmem_st8(tmem_lane + L::A_HI_COL + 8 * c, hi);
        tmem_st8(tmem_lane + L::A_LO_COL + 8 * c, lo);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    group_barrier(bar_id, L::GT);
    if (leader) {
      tc_fence_after();
      const uint32_t b0 = smem_u32(sB);
#pragma unroll
      for (int ks = 0; ks < L::KS; ++ks) {
        const uint64_t db = umma_desc(b0 + ks * 2 * L::B_LBO, L::B_LBO, 128);
        umma_tf32_ts(tmem_base, tmem_base + L::A_HI_COL + 8 * ks, db, idesc_wide, ks > 0);
        umma_tf32_ts(tmem_base + L::NP, tmem_base + L::A_LO_COL + 8 * ks, db, idesc_half, true);
      }
      umma_commit(mma_done);
      if (TMA && it + 2 < nitems) fence_proxy_async();
    }
    if (it + 2 < nitems) issue(s);
    if (leader) tma_store_wait_read();            // the stage is free again (previous item's store)
    if (fld == 0) load_j(tile + tstride, Jn);            // next tile's face Jacobians, behind the last use of Jf (see k_div_tc32)
    mbar_wait(mma_done, (uint32_t)it & 1u);
    tc_fence_after();
    group_barrier(bar_id, L::GT);                          // ... and every thread knows it (plain: all drained it)
    float* outp = static_cast<float*>(rows.out[fld]) + tile * L::TM * ND;
    float* stage = reinterpret_cast<float*>(stage_b + (TMA ? 0 : shift16(outp)));
#pragma unroll
    for (int q = 0; q < L::NQ; ++q) {
      if (L::NH > 1 && (q % L::NH) != half) continue;   // warp-uniform; q stays a compile-time constant
      float m[8], c[8];
      tmem_ld8(tmem_lane + 8 * q, m);
      tmem_ld8(tmem_lane + L::NP + 8 * q, c);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (8 * q + d < ND) stage[row * ND + 8 * q + d] = m[d] + c[d];
    }
    tc_fence_before();
    fence_proxy_async();
    group_barrier(bar_id, L::GT);
    if constexpr (TMA) {
      if (leader) {
        tma_store_2d(&maps.out[fld], stage_b, 0, (int)(tile * (L::TM / 4)));
        tma_store_commit();
      }
    } else {
      const long long e0 = tile * L::TM;
      plain_store<L::GT>(outp, stage_b, (int)(E - e0 < L::TM ? E - e0 : L::TM) * ND * 4, gtid);
    }
    if (++fld == nrows) { fld = 0; tile += tstride; }
  }
  if (leader) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ================================================================= DIV =====
// out[e,i] = sum_{r,j} D[r,i,j] w[r,e,j],  w[r,e,j] = sum_x J[x,r,e] u[x,e,j]:  N = 35 -> 48, K in three
// chunks (one per r) of 35 -> 40.  The thread that owns element e folds the Jacobian, splits w and
// writes chunk r into one of two A buffers in TMEM; the elected thread issues that chunk's MMAs while
// the group folds the next chunk.  TMEM columns of a group: [0,48) main, [48,96) corr,
// A buffer b: hi [96 + 80 b, +40), lo [136 + 80 b, +40).
template <int ND>
struct DivTC {
  static constexpr int TM = 128;
  static constexpr int KC = tc_pad(ND, 8), KS_C = KC / 8, NCHUNK = 3;   // per chunk: padded length (40), k-steps
  static constexpr int N = tc_pad(ND, 16), NB = 2 * N;  // 48; operator table rows [hi | lo] = 96
  static constexpr int WPG = 4, GT = 32 * WPG, NH = WPG / 4;
  static constexpr int GROUPS = ND <= 4 ? 8 : (ND <= 10 ? 5 : (ND <= 20 ? 3 : 2)), THREADS = GROUPS * GT;   // see GradTC
  static constexpr int B_LBO = NB * 16;                 // 1536
  static constexpr int B_BYTES = NCHUNK * (KC / 4) * B_LBO;   // 46 080
  static constexpr int U_SLAB = TM * ND;                // floats per x
  static constexpr int SLOT_BYTES = 3 * U_SLAB * 4;     // 53 760
  static constexpr int STAGE_BYTES = TM * ND * 4;       // 17 920
  static constexpr int NQ = (ND + 7) / 8;
  static constexpr int GROUP_BYTES = SLOT_BYTES + STAGE_BYTES;
  static constexpr int TMEM_COLS_PER_GROUP = GROUPS == 2 ? 256 : (GROUPS == 3 ? 160 : (GROUPS == 4 ? 128 : (GROUPS == 5 ? 96 : 64)));
  static constexpr int A_COL = NB, A_BUF = 2 * KC, A_LO = KC;
  static_assert(A_COL + 2 * A_BUF <= TMEM_COLS_PER_GROUP, "TMEM budget");
  static_assert(SLOT_BYTES % 128 == 0 && STAGE_BYTES % 128 == 0, "TMA alignment");
  static constexpr size_t SMEM = B_BYTES + (size_t)GROUPS * GROUP_BYTES + 512;   // + mbarriers, TMEM base
  // non-TMA variant (see GradTC): the three x-slabs and the stage carry kPlainPad bytes of slack
  static constexpr int U_SLAB_P = U_SLAB * 4 + kPlainPad;
  static constexpr int SLOT_P = 3 * U_SLAB_P, STAGE_P = STAGE_BYTES + kPlainPad;
  static constexpr int GROUP_P = SLOT_P + STAGE_P;
  static constexpr size_t SMEM_P = B_BYTES + (size_t)GROUPS * GROUP_P + 512;
  static_assert(SMEM_P <= 227 * 1024, "shared-memory budget");
};

struct DivTCMaps { CUtensorMap in, out; };

template <int ND, bool TMA = true>
__global__ void __launch_bounds__(DivTC<ND>::THREADS, 1)
k_div_tc32(const __grid_constant__ DivTCMaps maps, const float* __restrict__ Jg, const float* __restrict__ Dg,
           const float* __restrict__ ug, float* __restrict__ outg, long long E) {
  using L = DivTC<ND>;
  constexpr int GROUP_BYTES = TMA ? L::GROUP_BYTES : L::GROUP_P, SLOT_PITCH = TMA ? L::SLOT_BYTES : L::SLOT_P;
  constexpr int U_SLAB_B = TMA ? L::U_SLAB * 4 : L::U_SLAB_P;       // bytes between the x-slabs of the slot
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sB = smem_raw;
  unsigned char* groups = smem_raw + L::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(groups + (size_t)L::GROUPS * GROUP_BYTES);   // [group][full, mma0, mma1, mma2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * L::GROUPS);

  const int warp = uniform_warp_idx();
  const int gq = warp / L::WPG, wq = warp - gq * L::WPG;
  const int row = (wq & 3) * 32 + (threadIdx.x & 31);
  const int half = L::NH == 1 ? 0 : wq >> 2;
  const bool leader = wq == 0 && (threadIdx.x & 31) == 0;
  const int gtid = (int)threadIdx.x - gq * L::GT;

  if (threadIdx.x == 0) {
    for (int k = 0; k < 4 * L::GROUPS; ++k) mbar_init(&bars[k], (!TMA && k % 4 == 0) ? L::GT : 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // operator table [hi | lo]: row n = dof i, k = 40 r + j
  _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
  for (int idx = threadIdx.x; idx < L::N * L::NCHUNK * L::KC; idx += blockDim.x) {
    const int n = idx / (L::NCHUNK * L::KC), k = idx - n * (L::NCHUNK * L::KC);
    const int r = k / L::KC, j = k - r * L::KC;
    const float v = (n < ND && j < ND) ? Dg[(r * ND + n) * ND + j] : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const int off = (k >> 2) * L::B_LBO + (k & 3) * 4;
    *reinterpret_cast<uint32_t*>(sB + off + n * 16) = hi;
    *reinterpret_cast<uint32_t*>(sB + off + (L::N + n) * 16) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot + (uint32_t)gq * L::TMEM_COLS_PER_GROUP;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp & 3) * 32u << 16);

  unsigned char* gbase = groups + (size_t)gq * GROUP_BYTES;
  unsigned char* slot_b = gbase;
  unsigned char* stage_b = gbase + SLOT_PITCH;
  float* stage = reinterpret_cast<float*>(stage_b + (TMA ? 0 : shift16(outg)));   // tiles start at multiples of 512 ND bytes
  const float* sux[3];                                   // this thread's row in the three x-slabs
#pragma unroll
  for (int x = 0; x < 3; ++x)
    sux[x] = reinterpret_cast<const float*>(slot_b + x * U_SLAB_B + (TMA ? 0 : shift16(ug + (long long)x * E * ND))) + row * ND;
  uint64_t* full = &bars[4 * gq];
  uint64_t* mma_done = &bars[4 * gq + 1];                // one per chunk
  const int bar_id = 1 + gq;

  const long long ntiles = (E + L::TM - 1) / L::TM;
  const long long tile0 = (long long)blockIdx.x * L::GROUPS + gq, tstride = (long long)gridDim.x * L::GROUPS;
  constexpr uint32_t idesc_wide = umma_idesc_tf32(L::TM, L::NB), idesc_half = umma_idesc_tf32(L::TM, L::N);

  // TMA: the leader; plain: every thread of the group moves its share of the three x-slabs
  auto fetch = [&](long long tl) {
    if constexpr (TMA) {
      if (leader) {
        mbar_arrive_expect_tx(full, L::SLOT_BYTES);
        tma_load_3d(slot_b, &maps.in, 0, (int)(tl * (L::TM / 4)), 0, full);
      }
    } else {
      const long long e0 = tl * L::TM;
      const int nr = (int)(E - e0 < L::TM ? E - e0 : L::TM);
#pragma unroll
      for (int x = 0; x < 3; ++x)
        plain_load_async<L::GT>(slot_b + x * U_SLAB_B, ug + ((long long)x * E + e0) * ND, nr * ND * 4, gtid,
                                full, tl + 1 < ntiles);
      cp_async_arrive(full);
    }
  };
  if (tile0 < ntiles) fetch(tile0);
  float Jn[9];
  {
    const long long e = tile0 * L::TM + row;
#pragma unroll
    for (int xr = 0; xr < 9; ++xr) Jn[xr] = (tile0 < ntiles && e < E) ? __ldg(Jg + (long long)xr * E + e) : 0.f;
  }

  uint32_t it = 0;
  for (long long tile = tile0; tile < ntiles; tile += tstride, ++it) {
    float Jr[9];
#pragma unroll
    for (int xr = 0; xr < 9; ++xr) Jr[xr] = Jn[xr];
    mbar_wait(full, it & 1u);
#pragma unroll
    for (int r = 0; r < L::NCHUNK; ++r) {
      const int buf = r & 1;
      // chunk r - 2 used the same A buffer: its MMAs must have consumed it
      if (r == 2) { mbar_wait(&mma_done[0], it & 1u); tc_fence_after(); }
      const uint32_t a_hi = tmem_lane + L::A_COL + buf * L::A_BUF, a_lo = a_hi + L::A_LO;
#pragma unroll
      for (int c = 0; c < L::KC / 8; ++c) {
        if (L::NH > 1 && (c % L::NH) != half) continue;   // warp-uniform; c stays a compile-time constant
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = 8 * c + q;
          float w = 0.f;
          if (j < ND)
            w = fmaf(Jr[6 + r], sux[2][j], fmaf(Jr[3 + r], sux[1][j], Jr[r] * sux[0][j]));
          split_tf32(w, hi[q], lo[q]);
        }
        tmem_st8(a_hi + 8 * c, hi);
        tmem_st8(a_lo + 8 * c, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      group_barrier(bar_id, L::GT);
      if (leader) {
        tc_fence_after();
        const uint32_t b0 = smem_u32(sB) + r * (L::KC / 4) * L::B_LBO;
        const uint32_t ta = tmem_base + L::A_COL + buf * L::A_BUF;
#pragma unroll
        for (int ks = 0; ks < L::KS_C; ++ks) {
          const uint64_t db = umma_desc(b0 + ks * 2 * L::B_LBO, L::B_LBO, 128);
          umma_tf32_ts(tmem_base, ta + 8 * ks, db, idesc_wide, r > 0 || ks > 0);
          umma_tf32_ts(tmem_base + L::N, ta + L::A_LO + 8 * ks, db, idesc_half, true);
        }
        umma_commit(&mma_done[r]);
        if (TMA && r == L::NCHUNK - 1 && tile + tstride < ntiles) fence_proxy_async();
      }
      if (r == L::NCHUNK - 1) {
        // every thread is past its last read of the slot: fetch this group's next tile
        if (tile + tstride < ntiles) fetch(tile + tstride);
        if (leader) tma_store_wait_read();        // the stage is free again (previous tile's store)
      }
    }
    {
      // Jacobian of the next tile: issued only now, behind the last use of Jr.  (Issued at the top of the iteration,
      // the loads shared a scoreboard with the ones that produced Jr, and the first FMUL of the fold waited for them:
      // 12 % of the kernel's stall samples, profiles/r02_ncu_div_p4_f32_odd.txt.)  They land under the MMA wait.
      const long long tn = tile + tstride, e = tn * L::TM + row;
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jn[xr] = (tn < ntiles && e < E) ? __ldg(Jg + (long long)xr * E + e) : 0.f;
    }
    mbar_wait(&mma_done[1], it & 1u);
    mbar_wait(&mma_done[2], it & 1u);
    tc_fence_after();
    group_barrier(bar_id, L::GT);                          // stage free (leader waited above)
#pragma unroll
    for (int q = 0; q < L::NQ; ++q) {
      if (L::NH > 1 && (q % L::NH) != half) continue;   // warp-uniform; q stays a compile-time constant
      float m[8], c[8];
      tmem_ld8(tmem_lane + 8 * q, m);
      tmem_ld8(tmem_lane + L::N + 8 * q, c);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (8 * q + d < ND) stage[row * ND + 8 * q + d] = m[d] + c[d];
    }
    tc_fence_before();
    fence_proxy_async();
    group_barrier(bar_id, L::GT);
    if constexpr (TMA) {
      if (leader) {
        tma_store_2d(&maps.out, stage_b, 0, (int)(tile * (L::TM / 4)));
        tma_store_commit();
      }
    } else {
      const long long e0 = tile * L::TM;
      plain_store<L::GT>(outg + e0 * ND, stage_b, (int)(E - e0 < L::TM ? E - e0 : L::TM) * ND * 4, gtid);
    }
  }
  if (leader) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// tensor maps with 128-element boxes (element axis in quads of 140 floats = 560 B rows)
static bool map32_rows_box(CUtensorMap* tm, const void* base, long long E, int W, int box_quads) {
  const cuuint64_t dims[2] = {(cuuint64_t)(4 * W), (cuuint64_t)(E / 4)};
  const cuuint64_t strides[1] = {(cuuint64_t)(16 * W)};
  const cuuint32_t box[2] = {(cuuint32_t)(4 * W), (cuuint32_t)box_quads};
  return make_map32(tm, base, 2, dims, strides, box);
}
static bool map32_slabs_box(CUtensorMap* tm, const void* base, long long E, int W, int S, int box_quads) {
  const cuuint64_t dims[3] = {(cuuint64_t)(4 * W), (cuuint64_t)(E / 4), (cuuint64_t)S};
  const cuuint64_t strides[2] = {(cuuint64_t)(16 * W), (cuuint64_t)E * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)(4 * W), (cuuint32_t)box_quads, (cuuint32_t)S};
  return make_map32(tm, base, 3, dims, strides, box);
}

// order table of the compiled instantiations: tets p = 1..4
inline bool tc32_supported(int kind, int n_outer, int ni, int nj) {
  if (kind == FNSM_OP_GRAD || kind == FNSM_OP_DIV)
    return n_outer == 3 && ni == nj && (ni == 4 || ni == 10 || ni == 20 || ni == 35);
  return n_outer == 4 && ((ni == 4 && nj == 3) || (ni == 10 && nj == 6) || (ni == 20 && nj == 10) || (ni == 35 && nj == 15));
}

// Operands that qualify for tensor maps (E % 4 == 0, 16-byte aligned bases) take the TMA instantiation, everything
// else the plain-producer instantiation of the same kernel.
template <int ND>
static int launch_grad_tc32(const float* J, const float* D, const float* u, float* out, long long E,
                            const DevInfo& di, cudaStream_t st, bool force_plain) {
  using L = GradTC<ND>;
  if (E >= (1LL << 31) - L::TM) return FNSM_E_UNSUPPORTED;
  if (L::SMEM_P > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  const long long ntiles = (E + L::TM - 1) / L::TM;
  const long long need = (ntiles + L::GROUPS - 1) / L::GROUPS;
  const unsigned grid = (unsigned)(di.sms < need ? di.sms : need);
  GradTCMaps maps{};
  const bool tma = !force_plain && E % 4 == 0 && aligned16(u) && aligned16(out) &&
                   map32_rows_box(&maps.in, u, E, ND, L::TM / 4) && map32_slabs_box(&maps.out, out, E, ND, 3, L::TM / 4);
  if (tma) {
    if (int rc = set_smem(k_grad_tc32<ND, true>, L::SMEM)) return rc;
    k_grad_tc32<ND, true><<<grid, L::THREADS, L::SMEM, st>>>(maps, J, D, u, out, E);
  } else {
    if (int rc = set_smem(k_grad_tc32<ND, false>, L::SMEM_P)) return rc;
    k_grad_tc32<ND, false><<<grid, L::THREADS, L::SMEM_P, st>>>(maps, J, D, u, out, E);
  }
  return post_launch();
}

template <int ND, int NFD>
static int launch_lift_tc32(int kind, const float* J, const float* O, const OpmatRows& rows, int nrows, long long E,
                            const DevInfo& di, cudaStream_t st, bool force_plain) {
  using L = LiftTC<ND, NFD>;
  const bool fe = kind == FNSM_OP_LIFT_FE;
  if (E >= (1LL << 31) - L::TM) return FNSM_E_UNSUPPORTED;
  if (L::SMEM_P > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  LiftTCMaps maps{};
  bool tma = !force_plain && E % 4 == 0 && (fe || aligned16(J));
  for (int r = 0; r < nrows && tma; ++r)
    tma = aligned16(rows.field[r]) && aligned16(rows.out[r]) &&
          map32_slabs_box(&maps.in[r], rows.field[r], E, NFD, 4, L::TM / 4) &&
          map32_rows_box(&maps.out[r], rows.out[r], E, ND, L::TM / 4);
  const long long ntiles = (E + L::TM - 1) / L::TM;
  const long long need = (ntiles + L::GROUPS - 1) / L::GROUPS;
  const unsigned grid = (unsigned)(di.sms < need ? di.sms : need);
#define FNSM_LIFT_TC(FE_, TMA_, SMEM_)                                                          \
  do {                                                                                          \
    if (int rc = set_smem(k_lift_tc32<ND, NFD, FE_, TMA_>, SMEM_)) return rc;                   \
    k_lift_tc32<ND, NFD, FE_, TMA_><<<grid, L::THREADS, SMEM_, st>>>(maps, J, O, rows, nrows, E); \
  } while (0)
  if (fe) { if (tma) FNSM_LIFT_TC(true, true, L::SMEM); else FNSM_LIFT_TC(true, false, L::SMEM_P); }
  else    { if (tma) FNSM_LIFT_TC(false, true, L::SMEM); else FNSM_LIFT_TC(false, false, L::SMEM_P); }
#undef FNSM_LIFT_TC
  return post_launch();
}

template <int ND>
static int launch_div_tc32(const float* J, const float* D, const float* u, float* out, long long E,
                           const DevInfo& di, cudaStream_t st, bool force_plain) {
  using L = DivTC<ND>;
  if (E >= (1LL << 31) - L::TM) return FNSM_E_UNSUPPORTED;
  if (L::SMEM_P > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  const long long ntiles = (E + L::TM - 1) / L::TM;
  const long long need = (ntiles + L::GROUPS - 1) / L::GROUPS;
  const unsigned grid = (unsigned)(di.sms < need ? di.sms : need);
  DivTCMaps maps{};
  const bool tma = !force_plain && E % 4 == 0 && aligned16(u) && aligned16(out) &&
                   map32_slabs_box(&maps.in, u, E, ND, 3, L::TM / 4) && map32_rows_box(&maps.out, out, E, ND, L::TM / 4);
  if (tma) {
    if (int rc = set_smem(k_div_tc32<ND, true>, L::SMEM)) return rc;
    k_div_tc32<ND, true><<<grid, L::THREADS, L::SMEM, st>>>(maps, J, D, u, out, E);
  } else {
    if (int rc = set_smem(k_div_tc32<ND, false>, L::SMEM_P)) return rc;
    k_div_tc32<ND, false><<<grid, L::THREADS, L::SMEM_P, st>>>(maps, J, D, u, out, E);
  }
  return post_launch();
}

template <int ND, int NFD>
static int launch_tc32_order(int kind, const float* J, const float* O, const OpmatRows& rows, int nrows, long long E,
                             const DevInfo& di, cudaStream_t st, bool force_plain) {
  if (kind == FNSM_OP_LIFT_FE || kind == FNSM_OP_LIFT_EF)
    return launch_lift_tc32<ND, NFD>(kind, J, O, rows, nrows, E, di, st, force_plain);
  // every row of a batched grad / div is its own launch; lift walks its fields inside one launch
  for (int r = 0; r < nrows; ++r) {
    const float* u = static_cast<const float*>(rows.field[r]);
    float* out = static_cast<float*>(rows.out[r]);
    const int rc = kind == FNSM_OP_GRAD ? launch_grad_tc32<ND>(J, O, u, out, E, di, st, force_plain)
                                        : launch_div_tc32<ND>(J, O, u, out, E, di, st, force_plain);
    if (rc) return rc;
  }
  return FNSM_OK;
}

// force_plain (cfg->reserved[0] bit 0, the same debug bit as the DMMA kernels): take the plain-producer
// instantiation even when the operands qualify for TMA
static int launch_tc32(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                       int ni, long long E, const DevInfo& di, cudaStream_t st, bool force_plain = false) {
  const float* J = static_cast<const float*>(jac);
  const float* O = static_cast<const float*>(op);
  switch (ni) {
    case 4: return launch_tc32_order<4, 3>(kind, J, O, rows, nrows, E, di, st, force_plain);
    case 10: return launch_tc32_order<10, 6>(kind, J, O, rows, nrows, E, di, st, force_plain);
    case 20: return launch_tc32_order<20, 10>(kind, J, O, rows, nrows, E, di, st, force_plain);
    case 35: return launch_tc32_order<35, 15>(kind, J, O, rows, nrows, E, di, st, force_plain);
    default: return FNSM_E_UNSUPPORTED;
  }
}

}  // namespace fnsm
