// Variant 1 ("dmma") of the DG operator kernels for fp64, p = 4 tets.
//
// Every DG einsum is (tiny per-element scaling) o (constant matrix x element
// vector) -- SURVEY.md Appendix C.  The constant-matrix part runs on the FP64
// tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no FP64
// kind) with   M = 8 elements,  N = 8 output columns,  K = 4 contracted dofs.
//
// Structure of all three kernels ("load -> registers -> release -> DMMA stream"):
//   * one persistent CTA per SM, NW warps, NO producer warp and no CTA-wide
//     barrier in steady state.  Every warp owns one private shared-memory slot
//     and one mbarrier.  It waits for its slot, turns the slot into A fragments
//     held in registers (div folds the Jacobian, w = sum_x J[x,r,e] u[x,e,j];
//     lift scales by the face Jacobian; grad keeps u and applies J to the
//     accumulators afterwards -- the hoisting the reference's transforms do,
//     tuning/impls/xre_rij_ej_to_xei.py:104-117, xre_rij_xej_to_ei_v6.py:212-248,
//     ifj_fe_fej_to_ei_v3.py:94-105), and immediately re-arms the slot with the
//     TMA tensor copies (cp.async.bulk.tensor + mbarrier complete_tx) of its NEXT
//     work item, which then land while the warp issues its DMMAs.  Since only
//     the owning warp ever waits on a slot's barrier, phases cannot alias.
//   * TMA descriptors (tensor maps) view the element axis in PAIRS of elements
//     (70 / 30 doubles = 560 / 240 B rows, a multiple of 16 B), so one chunk is
//     ONE box per operand: 2 loads + 1 store per work item instead of 10 + 3
//     1-D copies -- the per-SM TMA unit was op-count bound with those -- and the
//     hardware zero-fills / clips the tail chunk.
//   * B fragments (the operator, zero padded in K to multiples of 4) sit in
//     shared memory in fragment order [tile][lane] -> conflict-free LDS.64,
//     each shared by the ME = 2 element tiles of a 16-element chunk.
//   * N = 35 output dofs are NOT padded to 40: dofs 0..31 take 4 DMMA column
//     tiles, dofs 32..34 are accumulated with DFMA from the same A registers
//     (each lane owns the k = t (mod 4) partial sums; two shuffles reduce).
//     That removes the 12.5 % padding waste of a fifth column tile.
//   * elements are permuted inside a chunk (el = 4*(g&3) + (g>>2) + 2m) so the
//     stride-35 / stride-15 rows read by a half-warp fall into distinct banks.
//
// Unaligned inputs (odd E, misaligned base) take the TMA = false instantiations: 1-D bulk copies (cp.async.bulk) of
// whole slabs, shifted by one double where a slab starts 8 (mod 16), completing on the same mbarrier; results leave
// through bulk stores from the stage (div_issue_bulk).  Small launches take the FS = true instantiations (k_div_dmma).
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudart)
#include <utility>
#include "common.cuh"
#include "opmat_simt.cuh"

namespace fnsm {

// Phase stamps for tools/timeline.cu (compiled out of the library): globaltimer at a few points of every warp's life,
// so that the fixed cost of a launch (prologue, operator staging, first load, tail) can be read off at small E.
#ifdef FNSM_TIMELINE
constexpr int kTlLaunches = 16, kTlSlots = 12;
__device__ unsigned long long fnsm_tl[kTlLaunches][160][16][kTlSlots];
__device__ unsigned fnsm_tl_ctr;
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define FNSM_TL_DECL __shared__ unsigned tl_launch_; \
  if (threadIdx.x == 0) tl_launch_ = (atomicAdd(&fnsm_tl_ctr, 1u) / gridDim.x) % kTlLaunches;
#define FNSM_TL(slot) do { if ((threadIdx.x & 31) == 0) fnsm_tl[tl_launch_][blockIdx.x][threadIdx.x >> 5][slot] = tl_now(); } while (0)
#define FNSM_TL_VAL(slot, v) do { if ((threadIdx.x & 31) == 0) fnsm_tl[tl_launch_][blockIdx.x][threadIdx.x >> 5][slot] = (v); } while (0)
#else
#define FNSM_TL_DECL
#define FNSM_TL(slot) do { } while (0)
#define FNSM_TL_VAL(slot, v) do { } while (0)
#endif

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
// orders this thread's earlier generic-proxy accesses to shared memory before
// later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA tensor copies global -> shared (completion on an mbarrier) and shared -> global
// (bulk async-group of the issuing thread).  `tm` points at a __grid_constant__ CUtensorMap.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      :: "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      :: "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               :: "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
               :: "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all earlier bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// 16-byte shared-memory load as a volatile asm statement: keeps its place among the (volatile) DMMAs
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double quad_sum(double v) {   // sum over the 4 lanes sharing g
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// 8-byte asynchronous global -> shared copy (always legal for fp64 operands) and its completion on an mbarrier:
// the plain (non-TMA) producer of the DMMA kernels.  A warp's lanes issue the copies of its next work item and arrive
// on the slot's barrier through their cp.async groups (barrier count 32 instead of 1), so the loads still land under
// the DMMA stream exactly like the TMA loads do.  (Round 1 copied synchronously: LDG -> STS -> arrive.)
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_rr(uint32_t dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
// the same with compile-time byte offsets folded into the instruction ([reg + imm] on both sides): one LDGSTS and
// nothing else per copy -- every extra instruction per copy is paid in full here (the kernels are issue bound)
template <int SOFF, int GOFF>
__device__ __forceinline__ void cp_async8_imm(uint32_t s, const double* g) {
  asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 8;" :: "r"(s), "l"(g), "n"(SOFF), "n"(GOFF) : "memory");
}
// lanes copy doubles lane + 32 q, q = 0 .. ceil(N / 32) - 1, of a contiguous run of N doubles (N known at compile time)
template <int N, int... Q>
__device__ __forceinline__ void cp_async_run_full(uint32_t s, const double* g, int lane, std::integer_sequence<int, Q...>) {
  // s and g already include the lane offset; only the last, partial, row of 32 needs a predicate
  (((Q + 1) * 32 <= N ? cp_async8_imm<Q * 256, Q * 256>(s, g)
                      : (lane + Q * 32 < N ? cp_async8_imm<Q * 256, Q * 256>(s, g) : (void)0)), ...);
}
template <int N>
__device__ __forceinline__ void cp_async_run(double* dst, const double* src, int n, int lane) {
  // n <= N valid doubles; n == N (every chunk but the last) takes the predicate-free form
  const uint32_t s = smem_u32(dst + lane);
  const double* g = src + lane;
  if (n == N) {
    cp_async_run_full<N>(s, g, lane, std::make_integer_sequence<int, (N + 31) / 32>{});
  } else {
    for (int k = lane; k < n; k += 32) cp_async8(dst + k, src + k);
  }
}
__device__ __forceinline__ void cp_async16(double* dst, const double* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
// Operator staging, first half: the raw operator (N doubles, row-major as the caller holds it) is copied into a
// scratch region of shared memory with coalesced cp.async -- every copy of the CTA in flight at once -- and the
// kernels then permute it into DMMA fragment order from shared memory.  Round 1 gathered the fragment order straight
// from global memory: 13 dependent-latency loads per thread on 8 sectors each, with all 148 CTAs hammering the same
// 29 KB of L2 -- 4.6 us of every launch (tools/timeline, profiles/r02_small_e.md), a fifth of the run time at
// E = 100 000.  The copies go out ahead of the first work-item loads (they gate everything); the caller then waits
// (stage_operator_raw_wait) and synchronises the CTA.
template <int N, int THREADS>
__device__ __forceinline__ void stage_operator_raw(double* scratch, const double* __restrict__ g) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    constexpr int N16 = N / 2;
#pragma unroll
    for (int q = 0; q < (N16 + THREADS - 1) / THREADS; ++q) {
      const int k = (int)threadIdx.x + q * THREADS;
      if (k < N16) cp_async16(scratch + 2 * k, g + 2 * k);
    }
    if ((N & 1) && threadIdx.x == 0) cp_async8(scratch + N - 1, g + N - 1);
  } else {
#pragma unroll
    for (int q = 0; q < (N + THREADS - 1) / THREADS; ++q) {
      const int k = (int)threadIdx.x + q * THREADS;
      if (k < N) cp_async8(scratch + k, g + k);
    }
  }
  // its own group: the wait below must not cover the work-item copies of the plain producers, which are issued
  // behind it and complete on their slot's mbarrier
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_operator_raw_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// 1-D bulk copy global -> shared with completion on an mbarrier: both addresses 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// ... and shared -> global (bulk async-group of the issuing thread, like the tensor stores)
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
// 1 if p is 8 (mod 16): a run of doubles that starts there is copied from one double earlier (and lands one double
// later in its shared-memory region, which has two doubles of headroom)
__device__ __forceinline__ int odd8(const void* p) { return (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1u); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// lets a kernel launched with programmatic stream serialization behind this one start as SMs free up (launch_k)
__device__ __forceinline__ void release_dependent_kernels() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Counterpart, executed by every thread right before it exits: a grid that was itself launched with programmatic
// stream serialization does not complete before the grid ahead of it has completed and flushed its writes, so
// "this grid is done" keeps implying "everything before it on the stream is done" (events, copies and later
// kernels order against the LAST kernel only).  No-op for a normally launched grid.
__device__ __forceinline__ void wait_for_previous_kernels() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// warp index as a value ptxas can prove warp-uniform (keeps TMA operands in uniform registers)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// one elected lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Start-up phase offset.  Warps w, w + 4, w + 8 share a sub-partition and its FP64 pipe; the pipe is
// arbitrated fairly, so warps that start together stay in lock step (all in their DMMA phase, then
// all converting / storing with the pipe idle: utilisation N S / (N S + Z) for N warps with S pipe
// cycles and Z other cycles per item).  Offsetting the first item of the k-th warp of a
// sub-partition by k * stagger cycles lets one warp's conversion and epilogue run under another's
// DMMA stream; the offset persists because equal sharing neither grows nor shrinks it.
__device__ __forceinline__ void stagger_start(int warp, int stagger) {
  const int rank = warp >> 2;
  if (stagger > 0 && rank > 0) {
    const long long t0 = clock64(), dt = (long long)rank * stagger;
    while (clock64() - t0 < dt) { }
  }
  __syncwarp();
}

// ------------------------------------------------------------ geometry -----
constexpr int kME = 2;            // element tiles (of 8) per chunk
constexpr int kCH = 8 * kME;      // elements per chunk / slot
constexpr int kNT = 4;            // DMMA column tiles: dofs 0..31
constexpr int kNL = 3;            // left-over dofs 32..34 (DFMA)

__device__ __forceinline__ int chunk_el(int g, int m) { return 4 * (g & 3) + (g >> 2) + 2 * m; }

constexpr int OUT_BLOCK = kCH * 35;   // doubles of one [16 elements][35 dofs] output block

// plain-path flush of one staged [kCH][35] block to out[e0 .. e0+kCH): coalesced stores
__device__ __forceinline__ void flush_plain(double* __restrict__ dst, const double* stage, long long e0,
                                            long long E, int lane) {
  const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
  for (int k = lane; k < ne * 35; k += 32) dst[k] = stage[k];
}

// tensor maps of one grad / div launch (element axis in pairs, see file header)
struct OpMaps { CUtensorMap in, jac, out; };
// lift: one Jacobian map, one input and one output map per field
struct LiftMaps { CUtensorMap jac; CUtensorMap in[8]; CUtensorMap out[8]; };

// launch flags (low bits of the kernels' `flags` argument)
constexpr int kFlagTma = 1;        // operands qualify for the TMA path
constexpr int kFlagNoLoad = 2;     // profiling aid: skip loads  (results invalid)
constexpr int kFlagNoStore = 4;    // profiling aid: skip stores (results invalid)
// The kernel is independent of the grid ahead of it on the stream (wave_3d_p4: the three einsums touch disjoint
// buffers): it loads right away and executes griddepcontrol.wait only before it exits.  Without the flag the wait
// comes before the first global access, so only the launch latency and the on-chip set-up overlap the previous tail.
constexpr int kFlagIndependent = 8;
constexpr int kDefaultStagger = 0; // cycles; bits 8.. of `flags` carry the start-up phase offset

// CTA-local dynamic work distribution: CTA b owns the items b, b + G, b + 2G, ...; its warps draw
// the next k from a shared-memory counter, so a sub-partition hosting fewer warps (NW not a
// multiple of 4) or a warp that was delayed simply takes fewer items.
struct WorkQueue {
  unsigned* ctr;
  long long nitems;
  __device__ __forceinline__ long long item_of(unsigned k) const {
    return (long long)blockIdx.x + (long long)k * gridDim.x;
  }
  __device__ __forceinline__ unsigned ticket(int lane) const {      // valid in lane 0 only
    return lane == 0 ? atomicAdd(ctr, 1u) : 0u;
  }
  __device__ __forceinline__ long long resolve(unsigned tk) const { return item_of(__shfl_sync(0xffffffffu, tk, 0)); }
  __device__ __forceinline__ long long take(int lane) const { return resolve(ticket(lane)); }
};

// ================================================================= DIV =====
// out[e,i] = sum_{r,j} D[r,i,j] * w[r,e,j],   w[r,e,j] = sum_x J[x,r,e] u[x,e,j]
// k-tiles ordered (jq, r): kt = 3*jq + r, k-in-tile t <-> j = 4*jq + t
// NX = 3: the divergence.  NX = 1: the same kernel without the x-sum, w[r,e,j] = J[r,e] u[e,j] -- the shared-operator
// family  se,sij,ej->ei  (reference test/test_codegen.py:34-88 "div components" / tuning/impls/re_rij_ej_to_ei*.py),
// which is  xse,sij,xej->ei  with |x| = 1: J(3,E), u(E,35).
template <int NX>
struct DivLayoutT {
  static constexpr int KT = 27;
  static constexpr int B_MAIN = KT * kNT * 32;                       // 3456
  static constexpr int B_LEFT = KT * 4 * 4;                          // [kt][t][3 dofs + pad]
  static constexpr int B_DOUBLES = B_MAIN + B_LEFT;                  // 3888
  static constexpr int U_SLAB = kCH * 35;                            // doubles per x
  static constexpr int SLOT_DOUBLES = NX * U_SLAB + 3 * NX * kCH;    // NX = 3: 1824 -> 14592 B
  static constexpr uint32_t SLOT_BYTES = SLOT_DOUBLES * 8;
  static constexpr int SLOT_DOUBLES_BULK = SLOT_DOUBLES + 2 * NX;    // TMA = false: two doubles of headroom per slab
};
using DivLayout = DivLayoutT<3>;

// ES (NX = 1 only): the geometric factors are laid out J(E, 3) -- "es,sij,ej->ei", reference
// examples/dg_wave_div.py:14 -- instead of J(3, E); the slot then holds them as [el][s]
//
// Producer for operands that do not qualify for tensor maps (odd E: every other slab / Jacobian row starts 8 (mod 16);
// or a view whose base does).  It lives in its own kernel instantiation (TMA = false): compiled into the TMA kernels
// as a cold branch -- inline or as a call -- it cost them 2-5 % (A/B on one box, profiles/r02_ab_dmma.md).
//   * a slab of a chunk is one contiguous run of 16 x 35 doubles whose offset inside the slab is a multiple of 16
//     bytes, so it travels as ONE 1-D bulk copy (cp.async.bulk); when the slab starts 8 (mod 16) the copy starts one
//     double early and is 16 bytes longer -- the slot's slab regions have two doubles of headroom (pitch U_SLAB + 2)
//     and the consumer reads slab x from s + x * pitch + sh[x].  Round 2's first version moved the slabs with 8-byte
//     cp.async from every lane, woven into the DMMA stream: 57 LDGSTS per lane and chunk against 3 bulk copies from
//     one lane -- in an issue-bound kernel that was the difference to the TMA instantiation (profiles/r02_cliff.md);
//   * the Jacobian entries (16 doubles per row) stay 8-byte cp.async: 5 per lane;
//   * the LAST chunk is copied element-wise: it may be partial, and a shifted bulk copy of it would read 8 bytes
//     past the end of the array.  (Inline: out of line -- a call -- the kernel lost 7 %; with byte counts computed
//     from the number of valid elements instead of this branch it lost 10 %.  ptxas' allocation of the hot loop
//     decides, not the instruction count.)
// The slot's barrier counts 33 arrivals: the 32 lanes through their cp.async groups + the elected lane's expect_tx
// (or plain) arrival.
template <int NX, bool ES>
__device__ __forceinline__ void div_issue_bulk(double* s, uint64_t* bar, const double* __restrict__ Jg,
                                               const double* __restrict__ ug, long long chunk, long long nchunks,
                                               long long E, int lane, const int (&sh)[NX]) {
  using L = DivLayoutT<NX>;
  constexpr int PITCH = L::U_SLAB + 2;
  const long long e0 = chunk * kCH;
  // rows of elements past E are not copied: they hold stale (finite or not) data, are computed -- the rows of a
  // DMMA tile are independent -- and never stored
  const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
  if (chunk + 1 < nchunks) {
    if (elect_one()) {
      fence_proxy_async();
      uint32_t bytes = 0;
#pragma unroll
      for (int x = 0; x < NX; ++x) bytes += L::U_SLAB * 8 + 16 * sh[x];
      mbar_arrive_expect_tx(bar, bytes);
#pragma unroll
      for (int x = 0; x < NX; ++x)
        bulk_load(s + x * PITCH, ug + ((long long)x * E + e0) * 35 - sh[x], L::U_SLAB * 8 + 16 * sh[x], bar);
    }
  } else {
#pragma unroll
    for (int x = 0; x < NX; ++x)
      cp_async_run<L::U_SLAB>(s + x * PITCH + sh[x], ug + ((long long)x * E + e0) * 35, ne * 35, lane);
    if (elect_one()) mbar_arrive(bar);
  }
  double* sJ = s + NX * PITCH;
  if (ES) {
#pragma unroll
    for (int q = 0; q < (3 * kCH + 31) / 32; ++q)
      if (lane + 32 * q < 3 * ne) cp_async8(sJ + lane + 32 * q, Jg + e0 * 3 + lane + 32 * q);
  } else {
    // k = lane + 32 q  <->  row xr = k / 16 = (lane >> 4) + 2 q, element el = lane & 15
    const int el = lane & (kCH - 1);
    const double* src = Jg + (long long)(lane >> 4) * E + e0 + el;
#pragma unroll
    for (int q = 0; q < (3 * NX * kCH + 31) / 32; ++q)
      if (lane + 32 * q < 3 * NX * kCH && el < ne) cp_async8(sJ + lane + 32 * q, src + (long long)(2 * q) * E);
  }
  cp_async_arrive_noinc(bar);
}

template <int NX, bool ES = false>
__device__ __forceinline__ void div_issue(double* s, uint64_t* bar, const OpMaps* maps, long long chunk) {
  using L = DivLayoutT<NX>;
  const long long e0 = chunk * kCH;
  if (elect_one()) {
    fence_proxy_async();
    mbar_arrive_expect_tx(bar, L::SLOT_BYTES);
    if (NX == 3) tma_load_3d(s, &maps->in, 0, (int)(chunk * (kCH / 2)), 0, bar);   // u[0..2][16 el][35]
    else         tma_load_2d(s, &maps->in, 0, (int)(chunk * (kCH / 2)), bar);      // u[16 el][35]
    if (ES) tma_load_2d(s + NX * L::U_SLAB, &maps->jac, 0, (int)(chunk * (kCH / 2)), bar);   // J[16 el][3]
    else    tma_load_2d(s + NX * L::U_SLAB, &maps->jac, (int)e0, 0, bar);      // J[3 NX][16 el]
  }
}

// (grad was tried with direct stores as well: its (x, element, dof-triple) store pattern costs 3.5x in time.)
// STAGED: results leave through a shared-memory stage + one TMA store (4.4 KB per warp); otherwise
// straight from the accumulator fragments with 8-byte streaming stores, which frees the shared
// memory for two more warps (19 KB per warp staged -> 10 warps, 14.6 KB direct -> 12 warps)
// DBG (profiling aid, results invalid): 1 = no conversion (A fragments = constants), 2 = no left-over
// DFMAs, 4 = no epilogue
// FS ("fast start", the instantiation small launches take): griddepcontrol.wait before the first global access (the
// launch then carries the programmatic-serialization attribute and its set-up hides behind the previous kernel's
// tail) and the operator staged through shared memory (stage_operator_raw).  Worth 4-5 us per launch; the hot loop
// ptxas builds around it is 1-2 % slower (register allocation), so larger launches (fast_start())
// keep the FS = false code, which is the round-1 kernel token for token.
template <int NW, bool STAGED, int DBG = 0, int NX = 3, bool ES = false, bool TMA = true, bool FS = false>
__global__ void __launch_bounds__(NW * 32, 1)
k_div_dmma(const __grid_constant__ OpMaps maps, const double* __restrict__ Jg, const double* __restrict__ Dg,
           const double* __restrict__ ug, double* __restrict__ outg, long long E, int flags) {
  using L = DivLayoutT<NX>;
  release_dependent_kernels();
  FNSM_TL_DECL
#ifdef FNSM_TIMELINE
  const unsigned long long tl_entry_ = tl_now();
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* sL = sB + L::B_MAIN;
  double* slots = sB + L::B_DOUBLES;
  constexpr int PITCH = L::U_SLAB + (TMA ? 0 : 2);                  // bulk producer: two doubles of headroom per slab
  constexpr int SLOT = L::SLOT_DOUBLES + (TMA ? 0 : 2 * NX);
  double* stages = slots + (size_t)NW * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (STAGED ? (size_t)NW * OUT_BLOCK : 0));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], TMA ? 1 : 33);   // bulk producer: 32 lanes + the elected one
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();
  FNSM_TL_VAL(0, tl_entry_);

  double* s = slots + (size_t)warp * SLOT;
  double* stage = stages + (size_t)warp * (STAGED ? OUT_BLOCK : 0);
  uint64_t* bar = &bars[warp];
  const double* sJ = s + NX * PITCH;
  const long long nchunks = (E + kCH - 1) / kCH;
  int sh[NX];                                                       // bulk producer: slab x sits at s + x PITCH + sh[x]
#pragma unroll
  for (int x = 0; x < NX; ++x) sh[x] = TMA ? 0 : odd8(ug + (long long)x * E * 35);
  auto issue = [&](long long chunk) {
    if constexpr (TMA) div_issue<NX, ES>(s, bar, &maps, chunk);
    else               div_issue_bulk<NX, ES>(s, bar, Jg, ug, chunk, nchunks, E, lane, sh);
  };
  const WorkQueue wq{work_ctr, nchunks};
  const int g = lane >> 2, t = lane & 3, tpad = t < 3 ? t : 2;

  constexpr bool tma = TMA;
  const bool dbg_noload = flags & kFlagNoLoad, dbg_nostore = flags & kFlagNoStore;
  if constexpr (FS) { if (!(flags & kFlagIndependent)) wait_for_previous_kernels(); }   // nothing global has been read yet
  FNSM_TL(1);
  // FS: tickets are drawn ONE item ahead (at the start of the iteration that needs them) instead of two.  With ~4
  // items per warp a depth of two fixes most of the assignment at launch, and a warp of a slower sub-partition ends
  // up with the last item while others idle (last warps 8-10 us behind the first, tools/timeline).
  long long cur = wq.take(lane), nxt = FS ? 0 : wq.take(lane);
  // The raw operator is staged through the slots of the last NSCR warps (there is no other free shared memory);
  // those warps issue their first loads once the tables are in fragment order.
  constexpr int RAW = 3 * 35 * 35;
  constexpr int NSCR = (RAW + SLOT - 1) / SLOT;
  constexpr bool kRawStage = FS && NSCR <= NW;
  static_assert(!FS || kRawStage, "fast start needs NSCR slots of scratch");
  const bool late = kRawStage && warp >= NW - NSCR;
  if constexpr (kRawStage) stage_operator_raw<RAW, NW * 32>(slots + (size_t)(NW - NSCR) * SLOT, Dg);
  if (cur < nchunks && !dbg_noload && !late) issue(cur);
  // operator tables are staged while the first TMA loads are in flight
  // main operator fragments, column tiles in pairs so that one LDS.128 feeds two of them:
  // sB[((kt*2 + p)*32 + lane)*2 + h] = D[r][8(2p+h)+g][4jq+t]
  // left-over dofs: sL[(kt*4 + t)*4 + d] = D[r][32+d][4jq+t]
  if constexpr (kRawStage) {
    const double* raw = slots + (size_t)(NW - NSCR) * SLOT;
    stage_operator_raw_wait();
    FNSM_TL(8);
    __syncthreads();
    FNSM_TL(9);
    // 128 table entries per k-tile and a CTA of a multiple of 128 threads: (h, lane, p) are fixed per thread and
    // only the k-tile moves, kt = kt0 + (NW / 4) q -- hardly any index arithmetic per entry (the generic loop
    // spent 1.5 us here, tools/timeline)
    static_assert(!FS || (NW % 4 == 0 && L::B_MAIN % (NW * 32) == 0), "table permutation assumes whole k-tiles per pass");
    {
      const int idx0 = (int)threadIdx.x & 127, kt0 = (int)threadIdx.x >> 7;
      const int h = idx0 & 1, ln = (idx0 >> 1) & 31, p = idx0 >> 6;
      const int gg = ln >> 2, tt = ln & 3;
      const double* src = raw + (8 * (2 * p + h) + gg) * 35 + tt;
#pragma unroll
      for (int q = 0; q < L::B_MAIN / (NW * 32); ++q) {
        const int kt = kt0 + (NW / 4) * q, jq = kt / 3, r = kt - 3 * jq;
        sB[kt * 128 + idx0] = (4 * jq + tt < 35) ? src[r * 1225 + 4 * jq] : 0.0;
      }
    }
    for (int idx = threadIdx.x; idx < L::B_LEFT; idx += blockDim.x) {
      const int d = idx & 3, t = (idx >> 2) & 3, kt = idx >> 4;
      const int jq = kt / 3, r = kt - 3 * jq, j = 4 * jq + t;
      sL[idx] = (d < kNL && j < 35) ? raw[(r * 35 + 32 + d) * 35 + j] : 0.0;
    }
    __syncthreads();
    if (cur < nchunks && !dbg_noload && late) issue(cur);
  } else {
    _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
    for (int idx = threadIdx.x; idx < L::B_MAIN; idx += blockDim.x) {
      const int h = idx & 1, ln = (idx >> 1) & 31, p = (idx >> 6) & 1, kt = idx >> 7;
      const int g = ln >> 2, t = ln & 3, jq = kt / 3, r = kt - 3 * jq;
      const int i = 8 * (2 * p + h) + g, j = 4 * jq + t;
      sB[idx] = (j < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.0;
    }
    _Pragma("unroll 4")
    for (int idx = threadIdx.x; idx < L::B_LEFT; idx += blockDim.x) {
      const int d = idx & 3, t = (idx >> 2) & 3, kt = idx >> 4;
      const int jq = kt / 3, r = kt - 3 * jq, j = 4 * jq + t;
      sL[idx] = (d < kNL && j < 35) ? Dg[(r * 35 + 32 + d) * 35 + j] : 0.0;
    }
    __syncthreads();
  }
  FNSM_TL(2);
  stagger_start(warp, flags >> 8);
  for (uint32_t n = 0; cur < nchunks; ++n) {
    if (!dbg_noload) mbar_wait(bar, n & 1u);
    unsigned tk = 0;
    if (FS) tk = wq.ticket(lane);                 // resolved behind the conversion
    if (n == 0) FNSM_TL(3);
    // ---- slot -> A fragments (Jacobian folded in) ----
    double a[kME][L::KT];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      if (DBG & 1) {
#pragma unroll
        for (int kt = 0; kt < L::KT; ++kt) a[m][kt] = (double)(kt + m) + (double)lane;
        continue;
      }
      const int el = chunk_el(g, m);
      double Jr[3 * NX];
#pragma unroll
      for (int xr = 0; xr < 3 * NX; ++xr) Jr[xr] = ES ? sJ[el * 3 + xr] : sJ[xr * kCH + el];
#pragma unroll
      for (int jq = 0; jq < 9; ++jq) {
        // k-slot j = 35 (jq = 8, t = 3) is padding: its operator entries are zero, so the lane reads j = 34 of its
        // own element there (finite whenever the element's data is) instead of paying a select per value
        const int j = jq == 8 ? 32 + tpad : 4 * jq + t;
        double ux[NX];
#pragma unroll
        for (int x = 0; x < NX; ++x) ux[x] = s[x * PITCH + sh[x] + el * 35 + j];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (NX == 3) a[m][3 * jq + r] = fma(Jr[2 * NX + r], ux[NX - 1], fma(Jr[NX + r], ux[NX > 1 ? 1 : 0], Jr[r] * ux[0]));
          else         a[m][3 * jq + r] = Jr[r] * ux[0];
        }
      }
    }
    __syncwarp();                                  // every lane is done reading the slot
    if (FS) nxt = wq.resolve(tk);
    if (nxt < nchunks && !dbg_noload) issue(nxt);
    if (!FS) tk = wq.ticket(lane);                // ticket after next; its latency hides under the DMMAs

    // ---- DMMA stream ----
    double acc[kME][kNT][2];
#pragma unroll
    for (int m = 0; m < kME; ++m)
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) { acc[m][nt][0] = 0.0; acc[m][nt][1] = 0.0; }
    // B fragments are fetched one k-tile ahead of their DMMAs (ptxas otherwise funnels every fragment
    // through one register pair and exposes the LDS latency to each pair of DMMAs)
    const uint32_t bB = smem_u32(sB) + lane * 16, bL = smem_u32(sL) + t * 32;
    {
    double2 bq[2][2];
    bq[0][0] = lds_v2(bB); bq[0][1] = lds_v2(bB + 512);
#pragma unroll
    for (int kt = 0; kt < L::KT; ++kt) {
      const int c = kt & 1, nx = c ^ 1;
      if (kt + 1 < L::KT) { bq[nx][0] = lds_v2(bB + (kt + 1) * 1024); bq[nx][1] = lds_v2(bB + (kt + 1) * 1024 + 512); }
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        dmma884(acc[m][0], a[m][kt], bq[c][0].x);
        dmma884(acc[m][1], a[m][kt], bq[c][0].y);
        dmma884(acc[m][2], a[m][kt], bq[c][1].x);
        dmma884(acc[m][3], a[m][kt], bq[c][1].y);
      }
    }
    }
    const long long e0 = cur * kCH;
    // ---- dofs 0..31 leave the registers first ... ----
    if (DBG & 4) {
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < kME; ++m)
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt) sum += acc[m][nt][0] + acc[m][nt][1];
      if (sum == 123.456) outg[0] = sum;
    } else if (STAGED) {
      if (lane == 0) tma_store_wait_read();          // previous block has left the stage
      __syncwarp();
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        double* o = stage + chunk_el(g, m) * 35;
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt) {          // rows of 35 doubles: odd elements are only 8-byte aligned
          o[8 * nt + 2 * t] = acc[m][nt][0];
          o[8 * nt + 2 * t + 1] = acc[m][nt][1];
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        const long long e = e0 + chunk_el(g, m);
        if (e < E && !dbg_nostore) {
          double* o = outg + e * 35;
#pragma unroll
          for (int nt = 0; nt < kNT; ++nt) {
            stg_stream(o + 8 * nt + 2 * t, acc[m][nt][0]);
            stg_stream(o + 8 * nt + 2 * t + 1, acc[m][nt][1]);
          }
        }
      }
    }
    // ---- ... then dofs 32..34 with DFMA from the same A registers.  A DFMA result is ready after ~30
    // cycles: with the 32 accumulator registers free again there is room for kLP partial sums per value
    // (6 kLP independent chains), which makes this tail issue bound (2 cycles per DFMA) instead of
    // latency bound (ptxas groups the DFMAs behind the DMMAs wherever they are written).
    // kLP swept on one box with the final kernel (bench.py div_p4, E = 4 M): 2: 1.072 ms, 3: 1.049, 4: 1.056,
    // 5: 1.028, 6: 1.029, 7: 1.032, 8: 1.041, 9: 1.047; 6 is also the best for the NX = 1 instantiation (se_p4)
#ifndef FNSM_KLP
#define FNSM_KLP 6
#endif
    constexpr int kLP = FNSM_KLP;
    double accL[kLP][kME][kNL];
#pragma unroll
    for (int p = 0; p < kLP; ++p)
#pragma unroll
      for (int m = 0; m < kME; ++m)
#pragma unroll
        for (int d = 0; d < kNL; ++d) accL[p][m][d] = 0.0;
    if (!(DBG & 2)) {
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
        const double2 l01 = lds_v2(bL + kt * 128), l2x = lds_v2(bL + kt * 128 + 16);
#pragma unroll
        for (int m = 0; m < kME; ++m) {
          accL[kt % kLP][m][0] = fma(a[m][kt], l01.x, accL[kt % kLP][m][0]);
          accL[kt % kLP][m][1] = fma(a[m][kt], l01.y, accL[kt % kLP][m][1]);
          accL[kt % kLP][m][2] = fma(a[m][kt], l2x.x, accL[kt % kLP][m][2]);
        }
      }
    }
    if (DBG & 4) {
      double sum = 0.0;
#pragma unroll
      for (int p = 0; p < kLP; ++p)
#pragma unroll
        for (int m = 0; m < kME; ++m)
#pragma unroll
          for (int d = 0; d < kNL; ++d) sum += accL[p][m][d];
      if (sum == 123.456) outg[1] = sum;
    } else {
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        double l[kNL];
#pragma unroll
        for (int d = 0; d < kNL; ++d) {
          double sacc = accL[0][m][d];
#pragma unroll
          for (int p = 1; p < kLP; ++p) sacc += accL[p][m][d];
          l[d] = quad_sum(sacc);
        }
        const double lv = t == 0 ? l[0] : (t == 1 ? l[1] : l[2]);
        if (STAGED) {
          if (t < kNL) stage[chunk_el(g, m) * 35 + 32 + t] = lv;
        } else {
          const long long e = e0 + chunk_el(g, m);
          if (e < E && !dbg_nostore && t < kNL) stg_stream(outg + e * 35 + 32 + t, lv);
        }
      }
      if (STAGED) {
        fence_proxy_async();
        __syncwarp();
        if (!dbg_nostore) {
          if (tma) {
            if (lane == 0) { tma_store_2d(&maps.out, stage, 0, (int)(cur * (kCH / 2))); tma_store_commit(); }
          } else {
            flush_plain(outg + e0 * 35, stage, e0, E, lane);
          }
        }
      }
    }
    if (n == 0) FNSM_TL(4);
    FNSM_TL_VAL(7, n + 1);
    cur = nxt;
    if (!FS) nxt = wq.resolve(tk);
  }
  FNSM_TL(5);
  if (lane == 0) tma_store_wait_all();
  FNSM_TL(6);
  wait_for_previous_kernels();   // kFlagIndependent: the only wait; otherwise a no-op (the grid ahead has completed)
}

// ================================================================ GRAD =====
// T[r][e][i] = sum_j D[r,i,j] u[e,j];  out[x,e,i] = sum_r J[x,r,e] T[r][e][i]
// k-tiles kt <-> j = 4*kt + t (9 tiles).  N = (dof, r) = 105 columns in 14 column
// tiles laid out so that a lane ends up with whole r-triples: lane t owns dofs
// 9t .. 9t+8 (t = 3: 27..34 plus one pad); its value v = 2*tile + h (h = column
// parity inside the tile, column c = 2t + h) is (dof 9t + v/3, r = v%3).  J is then
// applied in registers -- no shuffles, no left-over pass; 105/112 of the columns are useful.
struct GradLayout {
  static constexpr int KT = 9;
  static constexpr int NTILE = 14;
  static constexpr int B_DOUBLES = NTILE * KT * 32;                   // 4032
  static constexpr int U_SLAB = kCH * 35;
  static constexpr int SLOT_DOUBLES = U_SLAB + 9 * kCH;               // 704 -> 5632 B
  static constexpr uint32_t SLOT_BYTES = SLOT_DOUBLES * 8;
};

// bulk producer of a grad chunk (see div_issue_bulk): u as one 1-D bulk copy into a region with two doubles of
// headroom (data at s + shu), the 9 x 16 Jacobian entries by 8-byte cp.async from every lane, 33 arrivals
__device__ __forceinline__ void grad_issue_bulk(double* s, uint64_t* bar, const double* __restrict__ Jg,
                                                const double* __restrict__ ug, long long chunk, long long nchunks,
                                                long long E, int lane, int shu) {
  using L = GradLayout;
  const long long e0 = chunk * kCH;
  const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
  if (chunk + 1 < nchunks) {
    if (elect_one()) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, L::U_SLAB * 8 + 16 * shu);
      bulk_load(s, ug + e0 * 35 - shu, L::U_SLAB * 8 + 16 * shu, bar);
    }
  } else {                                   // last chunk: element-wise (partial; end of the array)
    cp_async_run<L::U_SLAB>(s + shu, ug + e0 * 35, ne * 35, lane);
    if (elect_one()) mbar_arrive(bar);
  }
  {
    const int el = lane & (kCH - 1);
    const double* src = Jg + (long long)(lane >> 4) * E + e0 + el;
#pragma unroll
    for (int q = 0; q < (9 * kCH + 31) / 32; ++q)
      if (lane + 32 * q < 9 * kCH && el < ne) cp_async8(s + L::U_SLAB + 2 + lane + 32 * q, src + (long long)(2 * q) * E);
  }
  cp_async_arrive_noinc(bar);
}

// TMA instantiation: `tma_ok` = the launch flag kFlagTma.  With it clear the slot is filled synchronously (LDG -> STS ->
// one arrival: round 1's plain path).  The library's launchers never clear it -- operands that do not qualify for
// tensor maps go to the TMA = false instantiation (bulk producer) -- but the branch stays: with it compiled out,
// ptxas allocates / schedules the hot loop of the grad and lift kernels 1-3 % slower (A/B of both builds on one box,
// profiles/r02_ab_dmma.md).  It costs nothing at run time (one uniform predicate per work item).
__device__ __forceinline__ void grad_issue(double* s, uint64_t* bar, const OpMaps* maps,
                                           const double* __restrict__ Jg, const double* __restrict__ ug,
                                           long long chunk, long long E, bool tma_ok, int lane) {
  using L = GradLayout;
  const long long e0 = chunk * kCH;
  if (tma_ok) {
    if (elect_one()) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, L::SLOT_BYTES);
      tma_load_2d(s, &maps->in, 0, (int)(chunk * (kCH / 2)), bar);               // u[16 el][35]
      tma_load_2d(s + L::U_SLAB, &maps->jac, (int)e0, 0, bar);                   // J[9][16 el]
    }
  } else {
    const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
    for (int k = lane; k < L::U_SLAB; k += 32) s[k] = (k < ne * 35) ? ug[e0 * 35 + k] : 0.0;
    for (int k = lane; k < 9 * kCH; k += 32) {
      const int xr = k / kCH, el = k - xr * kCH;
      s[L::U_SLAB + k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : 0.0;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  }
}

// one group of NTG column tiles (starting at tile T0): DMMAs, then J applied to the
// NTG*2/3 complete (dof, r)-triples this lane now holds, results staged in shared memory
// BULK (TMA = false instantiation): the x-slabs of the stage are OUT_BLOCK + 2 apart and slab x holds its data from
// so[x] on -- 1 where the slab's run in global memory starts 8 (mod 16), so that the bulk store (which skips the first
// and the last double then) finds a 16-byte aligned source
template <int T0, int NTG, bool BULK = false>
__device__ __forceinline__ void grad_group(const double* __restrict__ sB, const double (&a)[kME][GradLayout::KT],
                                           const double (&Jr)[kME][9], double* stage, int g, int t, int lane,
                                           const int* so = nullptr) {
  using L = GradLayout;
  double acc[kME][NTG][2];
#pragma unroll
  for (int m = 0; m < kME; ++m)
#pragma unroll
    for (int jt = 0; jt < NTG; ++jt) { acc[m][jt][0] = 0.0; acc[m][jt][1] = 0.0; }
#pragma unroll
  for (int kt = 0; kt < L::KT; ++kt) {
#pragma unroll
    for (int jt = 0; jt < NTG; ++jt) {
      const double b = sB[((T0 + jt) * L::KT + kt) * 32 + lane];
#pragma unroll
      for (int m = 0; m < kME; ++m) dmma884(acc[m][jt], a[m][kt], b);
    }
  }
  if (T0 == 0) {                             // first write of this chunk into the stage:
    if (lane == 0) tma_store_wait_read();    // the previous chunk's bulk store must have read it out
    __syncwarp();
  }
  constexpr int V0 = 2 * T0;                 // first value index of this group (multiple of 3 by construction)
  static_assert(V0 % 3 == 0, "groups must start on a triple boundary");
  constexpr int NTRI = (2 * NTG) / 3;        // complete triples in the group
#pragma unroll
  for (int m = 0; m < kME; ++m) {
    double* o = stage + chunk_el(g, m) * 35 + 9 * t + V0 / 3;
#pragma unroll
    for (int q = 0; q < NTRI; ++q) {
      const double T0v = acc[m][(3 * q) >> 1][(3 * q) & 1];
      const double T1v = acc[m][(3 * q + 1) >> 1][(3 * q + 1) & 1];
      const double T2v = acc[m][(3 * q + 2) >> 1][(3 * q + 2) & 1];
      if (V0 / 3 + q < 8 || t < 3) {          // dof 9t + 8 exists only for t < 3
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const double v = fma(Jr[m][3 * x + 2], T2v, fma(Jr[m][3 * x + 1], T1v, Jr[m][3 * x] * T0v));
          if constexpr (BULK) o[x * (OUT_BLOCK + 2) + so[x] + q] = v;
          else                o[x * OUT_BLOCK + q] = v;
        }
      }
    }
  }
}

template <int NW, bool TMA = true, bool FS = false>      // FS: see k_div_dmma
__global__ void __launch_bounds__(NW * 32, 1)
k_grad_dmma(const __grid_constant__ OpMaps maps, const double* __restrict__ Jg, const double* __restrict__ Dg,
            const double* __restrict__ ug, double* __restrict__ outg, long long E, int flags) {
  using L = GradLayout;
  release_dependent_kernels();
  FNSM_TL_DECL
#ifdef FNSM_TIMELINE
  const unsigned long long tl_entry_ = tl_now();
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* slots = sB + L::B_DOUBLES;
  constexpr int PAD = TMA ? 0 : 2;                                  // bulk producer / stores: headroom per slab
  constexpr int SLOT = L::SLOT_DOUBLES + PAD, SPITCH = OUT_BLOCK + PAD;
  double* stages = slots + (size_t)NW * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)NW * 3 * SPITCH);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], TMA ? 1 : 33);   // bulk producer: 32 lanes + the elected one
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();
  FNSM_TL_VAL(0, tl_entry_);

  double* s = slots + (size_t)warp * SLOT;
  double* stage = stages + (size_t)warp * 3 * SPITCH;         // [x][16][35]
  uint64_t* bar = &bars[warp];
  const double* sJ = s + L::U_SLAB + PAD;
  // bulk instantiation: u sits at s + shu; slab x of the stage holds its data from so[x] on (see grad_group)
  const int shu = TMA ? 0 : odd8(ug);
  int so[3];
#pragma unroll
  for (int x = 0; x < 3; ++x) so[x] = TMA ? 0 : odd8(outg + (long long)x * E * 35);
  const long long nchunks = (E + kCH - 1) / kCH;
  const WorkQueue wq{work_ctr, nchunks};
  const int g = lane >> 2, t = lane & 3, tpad = t < 3 ? t : 2;

  const bool tma = TMA && (flags & kFlagTma);
  const bool dbg_noload = flags & kFlagNoLoad, dbg_nostore = flags & kFlagNoStore;
  if constexpr (FS) { if (!(flags & kFlagIndependent)) wait_for_previous_kernels(); }   // nothing global has been read yet
  FNSM_TL(1);
  long long cur = wq.take(lane), nxt = FS ? 0 : wq.take(lane);     // FS: tickets one item ahead, see k_div_dmma
  // fast start: raw D goes into the (still unused) output stages, ahead of the first work-item loads
  static_assert((size_t)NW * 3 * OUT_BLOCK >= 3 * 35 * 35, "raw operator does not fit the stage area");
  if constexpr (FS) stage_operator_raw<3 * 35 * 35, NW * 32>(stages, Dg);
  if (cur < nchunks && !dbg_noload) {
    if constexpr (TMA) grad_issue(s, bar, &maps, Jg, ug, cur, E, tma, lane);
    else               grad_issue_bulk(s, bar, Jg, ug, cur, nchunks, E, lane, shu);
  }
  if constexpr (FS) {
    // sB[(tile*KT + kt)*32 + lane]: B[k = t][n = c], c = lane>>2 -> value v = 2*tile + (c&1) of lane c>>1
    stage_operator_raw_wait();
    FNSM_TL(8);
    __syncthreads();
    FNSM_TL(9);
    // a warp permutes units of (column tile, three k-tiles): the (dof, r) decode is done once per unit
    const int c = lane >> 2, tt = lane & 3;
#pragma unroll
    for (int q = 0; q < (L::NTILE * 3 + NW - 1) / NW; ++q) {
      const int u = warp + NW * q;
      if (u < L::NTILE * 3) {
        const int tile = u / 3, ktb = 3 * (u - 3 * tile);
        const int v = 2 * tile + (c & 1), dv = v / 3, i = 9 * (c >> 1) + dv, r = v - 3 * dv;
        const bool ok = v < 27 && i < 35;
        const double* src = stages + (r * 35 + i) * 35 + tt;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int kt = ktb + kk;
          sB[(tile * L::KT + kt) * 32 + lane] = (ok && 4 * kt + tt < 35) ? src[4 * kt] : 0.0;
        }
      }
    }
  } else {
    // operator tables are staged while the first TMA loads are in flight
    // sB[(tile*KT + kt)*32 + lane]: B[k = t][n = c], c = lane>>2 -> value v = 2*tile + (c&1) of lane c>>1
    _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
    for (int idx = threadIdx.x; idx < L::B_DOUBLES; idx += blockDim.x) {
      const int ln = idx & 31, kt = (idx >> 5) % L::KT, tile = (idx >> 5) / L::KT;
      const int c = ln >> 2, t = ln & 3;
      const int v = 2 * tile + (c & 1), i = 9 * (c >> 1) + v / 3, r = v % 3, j = 4 * kt + t;
      sB[idx] = (v < 27 && i < 35 && j < 35) ? Dg[(r * 35 + i) * 35 + j] : 0.0;
    }
  }
  __syncthreads();
  FNSM_TL(2);
  stagger_start(warp, flags >> 8);
  for (uint32_t n = 0; cur < nchunks; ++n) {
    if (!dbg_noload) mbar_wait(bar, n & 1u);
    unsigned tk = 0;
    if (FS) tk = wq.ticket(lane);
    if (n == 0) FNSM_TL(3);
    double a[kME][L::KT];
    double Jr[kME][9];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt)       // k-slot j = 35 is padding (zero operator entries): reads j = 34
        a[m][kt] = s[shu + el * 35 + (kt == 8 ? 32 + tpad : 4 * kt + t)];
#pragma unroll
      for (int xr = 0; xr < 9; ++xr) Jr[m][xr] = sJ[xr * kCH + el];
    }
    __syncwarp();
    if (FS) nxt = wq.resolve(tk);
    if (nxt < nchunks && !dbg_noload) {
      if constexpr (TMA) grad_issue(s, bar, &maps, Jg, ug, nxt, E, tma, lane);
      else               grad_issue_bulk(s, bar, Jg, ug, nxt, nchunks, E, lane, shu);
    }
    if (!FS) tk = wq.ticket(lane);

    const long long e0 = cur * kCH;
    grad_group<0, 3, !TMA>(sB, a, Jr, stage, g, t, lane, so);
    grad_group<3, 3, !TMA>(sB, a, Jr, stage, g, t, lane, so);
    grad_group<6, 3, !TMA>(sB, a, Jr, stage, g, t, lane, so);
    grad_group<9, 3, !TMA>(sB, a, Jr, stage, g, t, lane, so);
    grad_group<12, 2, !TMA>(sB, a, Jr, stage, g, t, lane, so);
    fence_proxy_async();
    __syncwarp();
    if (!dbg_nostore) {
      if constexpr (TMA) {
        if (tma) {
          if (lane == 0) { tma_store_3d(&maps.out, stage, 0, (int)(cur * (kCH / 2)), 0); tma_store_commit(); }
        } else {
#pragma unroll
          for (int x = 0; x < 3; ++x)
            flush_plain(outg + ((long long)x * E + e0) * 35, stage + x * OUT_BLOCK, e0, E, lane);
        }
      } else if (cur + 1 < nchunks) {
        // whole chunk: one bulk store per x-slab; a run that starts 8 (mod 16) leaves without its first and last
        // double, which two lanes store directly
        if (lane == 0) {
#pragma unroll
          for (int x = 0; x < 3; ++x)
            bulk_store(outg + ((long long)x * E + e0) * 35 + so[x], stage + x * SPITCH + 2 * so[x], OUT_BLOCK * 8 - 16 * so[x]);
          tma_store_commit();
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          if (so[x] && lane >= 1 && lane < 3) {
            const int k = lane == 1 ? 0 : OUT_BLOCK - 1;
            outg[((long long)x * E + e0) * 35 + k] = stage[x * SPITCH + 1 + k];
          }
        }
      } else {
#pragma unroll
        for (int x = 0; x < 3; ++x)
          flush_plain(outg + ((long long)x * E + e0) * 35, stage + x * SPITCH + so[x], e0, E, lane);
      }
    }
    if (n == 0) FNSM_TL(4);
    FNSM_TL_VAL(7, n + 1);
    cur = nxt;
    if (!FS) nxt = wq.resolve(tk);
  }
  FNSM_TL(5);
  if (lane == 0) tma_store_wait_all();
  FNSM_TL(6);
  wait_for_previous_kernels();   // kFlagIndependent: the only wait; otherwise a no-op (the grid ahead has completed)
}

// ================================================================ LIFT =====
// out_k[e,i] = sum_{f,j} Op(f,i,j) * Jf(e,f) * v_k[f,e,j];  K = (f,j) = 60 = 15 k-tiles
// work item = (chunk, field): a slot holds one field of one chunk
// position of the operator fragment of (k-tile, column tile, lane) in the table.  PAIRED: column tiles in pairs, so
// that one LDS.128 feeds two of them (as the divergence kernel does): 30 instead of 60 fragment loads per item.  Taken
// by the TMA = false instantiations only (E = 4 000 001: 73.5 -> 75.0 % of roofline); the same change makes the TMA
// instantiation 3 % slower (2.305 -> 2.370 ms at E = 4 M, A/B on one box) -- ptxas' allocation again.
template <bool PAIRED>
__device__ __forceinline__ int lift_b_pos(int kt, int nt, int ln) {
  return PAIRED ? ((kt * 2 + (nt >> 1)) * 32 + ln) * 2 + (nt & 1) : (kt * kNT + nt) * 32 + ln;
}
struct LiftLayout {
  static constexpr int KT = 15;
  static constexpr int B_MAIN = KT * kNT * 32;                        // 1920
  static constexpr int B_LEFT = KT * 4 * 4;                           // 240
  static constexpr int B_DOUBLES = B_MAIN + B_LEFT;
  static constexpr int V_SLAB = kCH * 15;                             // doubles per face
  static constexpr int SLOT_DOUBLES = 4 * V_SLAB + 4 * kCH;           // 1024 -> 8192 B
  static constexpr uint32_t SLOT_BYTES = SLOT_DOUBLES * 8;
};

// bulk producer of a lift item (see div_issue_bulk): the four face slabs (16 x 15 doubles each) as 1-D bulk copies
// into regions with two doubles of headroom (slab f at s + f (V_SLAB + 2) + ((shm >> f) & 1)), the 4 x 16 face
// Jacobians by 8-byte cp.async from every lane, 33 arrivals
__device__ __forceinline__ int lift_shift_mask(const double* vg, long long E) {
  int m = 0;
#pragma unroll
  for (int f = 0; f < 4; ++f) m |= odd8(vg + (long long)f * E * 15) << f;
  return m;
}
template <bool FE>
__device__ __forceinline__ void lift_issue_bulk(double* s, uint64_t* bar, const double* __restrict__ Jg,
                                                const double* __restrict__ vg, long long chunk, long long nchunks,
                                                long long E, int lane, int shm) {
  using L = LiftLayout;
  constexpr int P = L::V_SLAB + 2;
  const long long e0 = chunk * kCH;
  const int ne = (int)((E - e0 < kCH) ? (E - e0) : kCH);
  if (chunk + 1 < nchunks) {
    if (elect_one()) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bar, 4 * L::V_SLAB * 8 + 16 * __popc(shm));
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int sh = (shm >> f) & 1;
        bulk_load(s + f * P, vg + ((long long)f * E + e0) * 15 - sh, L::V_SLAB * 8 + 16 * sh, bar);
      }
    }
  } else {                                   // last chunk: element-wise (partial; end of the arrays)
#pragma unroll
    for (int f = 0; f < 4; ++f)
      cp_async_run<L::V_SLAB>(s + f * P + ((shm >> f) & 1), vg + ((long long)f * E + e0) * 15, ne * 15, lane);
    if (elect_one()) mbar_arrive(bar);
  }
#pragma unroll
  for (int q = 0; q < (4 * kCH) / 32; ++q) {
    const int k = lane + 32 * q;
    if (FE) { const int f = k / kCH, el = k - f * kCH; if (el < ne) cp_async8(s + 4 * P + k, Jg + (long long)f * E + e0 + el); }
    else    { const int el = k / 4; if (el < ne) cp_async8(s + 4 * P + k, Jg + e0 * 4 + k); }
  }
  cp_async_arrive_noinc(bar);
}

template <bool FE>
__device__ __forceinline__ void lift_issue(double* s, uint64_t* bar, const CUtensorMap* map_v,
                                           const CUtensorMap* map_j, const double* __restrict__ Jg,
                                           const double* __restrict__ vg, long long chunk, long long E,
                                           bool tma_ok, int lane) {
  using L = LiftLayout;
  const long long e0 = chunk * kCH;
  {
    if (tma_ok) {
      if (elect_one()) {
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, L::SLOT_BYTES);
        tma_load_3d(s, map_v, 0, (int)(chunk * (kCH / 2)), 0, bar);                // v[4][16 el][15]
        if (FE) tma_load_2d(s + 4 * L::V_SLAB, map_j, (int)e0, 0, bar);            // Jface[4][16 el]
        else    tma_load_2d(s + 4 * L::V_SLAB, map_j, 0, (int)(chunk * (kCH / 2)), bar);   // J[16 el][4]
      }
    } else {                                                                       // see grad_issue
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
      if (lane == 0) mbar_arrive(bar);
    }
  }
}

// FS ("fast start", see k_div_dmma) also switches the queue from whole chunks to single items (chunk, field): at
// E = 100 000 a chunk-granular queue leaves 2.6 units per warp -- a 3 : 2 spread between warps and a 10 us tail
// (tools/timeline); large launches keep the coarse queue, whose loop carries less state.
template <int NW, bool FE, bool TMA = true, bool FS = false>
__global__ void __launch_bounds__(NW * 32, 1)
k_lift_dmma(const __grid_constant__ LiftMaps maps, const double* __restrict__ Jg,
            const double* __restrict__ Og, const __grid_constant__ OpmatRows rows, int nrows, long long E, int flags) {
  using L = LiftLayout;
  release_dependent_kernels();
  FNSM_TL_DECL
#ifdef FNSM_TIMELINE
  const unsigned long long tl_entry_ = tl_now();
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sB = reinterpret_cast<double*>(smem_raw);
  double* sL = sB + L::B_MAIN;
  double* slots = sB + L::B_DOUBLES;
  constexpr int PAD = TMA ? 0 : 2;                                  // bulk producer / stores: headroom per slab
  constexpr int P = L::V_SLAB + PAD, SLOT = L::SLOT_DOUBLES + 4 * PAD, SPITCH = OUT_BLOCK + PAD;
  double* stages = slots + (size_t)NW * SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)NW * SPITCH);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  unsigned* work_ctr = reinterpret_cast<unsigned*>(bars + NW);
  if (threadIdx.x == 0) {
    for (int w = 0; w < NW; ++w) mbar_init(&bars[w], TMA ? 1 : 33);   // bulk producer: 32 lanes + the elected one
    *work_ctr = 0u;
    mbar_fence_init();
  }
  __syncthreads();
  FNSM_TL_VAL(0, tl_entry_);

  double* s = slots + (size_t)warp * SLOT;
  double* stage = stages + (size_t)warp * SPITCH;
  uint64_t* bar = &bars[warp];
  const double* sJ = s + 4 * P;
  const long long nchunks = (E + kCH - 1) / kCH;
  // work unit of the queue: a chunk (the warp walks its fields), or with FS one item, id = chunk * nrows + field
  // (< 2^30: division by nrows <= 8 as a multiplication by ceil(2^32 / nrows), exact in that range)
  const WorkQueue wq{work_ctr, FS ? nchunks * nrows : nchunks};
  // (nrows = 1: the multiplier 2^32 does not fit, the quotient is the id itself)
  const unsigned magic = nrows > 1 ? (unsigned)((0x100000000ull + (unsigned)nrows - 1) / (unsigned)nrows) : 0u;
  const int g = lane >> 2, t = lane & 3;

  const bool tma = TMA && (flags & kFlagTma);
  if constexpr (FS) { if (!(flags & kFlagIndependent)) wait_for_previous_kernels(); }   // nothing global has been read yet
  FNSM_TL(1);
  // loop state: chunk and field of the current item, queue position after it
  long long cur = wq.take(lane), nxt = FS ? 0 : wq.take(lane);     // FS: tickets one item ahead, see k_div_dmma
  int fld = 0;
  if (FS) {
    const unsigned q = nrows > 1 ? __umulhi((unsigned)cur, magic) : (unsigned)cur;
    fld = (int)((unsigned)cur - q * (unsigned)nrows);
    cur = q;
  }
  // fast start: the raw operator goes into the (still unused) output stages, ahead of the first work-item loads
  static_assert((size_t)NW * OUT_BLOCK >= 35 * 4 * 15, "raw operator does not fit the stage area");
  if constexpr (FS) stage_operator_raw<35 * 4 * 15, NW * 32>(stages, Og);
  // bulk instantiation: shift bits of the four face slabs of the item in the slot (they depend on the field's base)
  int shm = 0;
  if (cur < nchunks) {
    const double* vg = static_cast<const double*>(rows.field[fld]);
    if constexpr (TMA) {
      lift_issue<FE>(s, bar, &maps.in[fld], &maps.jac, Jg, vg, cur, E, tma, lane);
    } else {
      shm = lift_shift_mask(vg, E);
      lift_issue_bulk<FE>(s, bar, Jg, vg, cur, nchunks, E, lane, shm);
    }
  }
  if constexpr (FS) {
    // sB[(kt*4 + nt)*32 + lane] = Op(f, 8nt+g, j),  k = 4kt+t = 15 f + j
    stage_operator_raw_wait();
    FNSM_TL(8);
    __syncthreads();
    FNSM_TL(9);
#pragma unroll
    for (int q = 0; q < (L::B_MAIN + NW * 32 - 1) / (NW * 32); ++q) {
      const int idx = (int)threadIdx.x + q * NW * 32;
      if (idx < L::B_MAIN) {
        const int ln = idx & 31, nt = (idx >> 5) % kNT, kt = (idx >> 5) / kNT;
        const int g = ln >> 2, t = ln & 3, k = 4 * kt + t, f = k / 15, j = k - 15 * f;
        const int i = 8 * nt + g;
        sB[TMA ? idx : lift_b_pos<true>(kt, nt, ln)] = FE ? stages[(i * 4 + f) * 15 + j] : stages[(f * 35 + i) * 15 + j];
      }
    }
    for (int idx = threadIdx.x; idx < L::B_LEFT; idx += blockDim.x) {
      const int d = idx & 3, t = (idx >> 2) & 3, kt = idx >> 4;
      const int k = 4 * kt + t, f = k / 15, j = k - 15 * f, i = 32 + d;
      double v = 0.0;
      if (d < kNL) v = FE ? stages[(i * 4 + f) * 15 + j] : stages[(f * 35 + i) * 15 + j];
      sL[idx] = v;
    }
  } else {
    // operator tables are staged while the first TMA loads are in flight
    // sB[(kt*4 + nt)*32 + lane] = Op(f, 8nt+g, j),  k = 4kt+t = 15 f + j
    _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
    for (int idx = threadIdx.x; idx < L::B_MAIN; idx += blockDim.x) {
      const int ln = idx & 31, nt = (idx >> 5) % kNT, kt = (idx >> 5) / kNT;
      const int g = ln >> 2, t = ln & 3, k = 4 * kt + t, f = k / 15, j = k - 15 * f;
      const int i = 8 * nt + g;
      sB[TMA ? idx : lift_b_pos<true>(kt, nt, ln)] = FE ? Og[(i * 4 + f) * 15 + j] : Og[(f * 35 + i) * 15 + j];
    }
    _Pragma("unroll 4")   // independent operator loads in flight (matters at small E)
    for (int idx = threadIdx.x; idx < L::B_LEFT; idx += blockDim.x) {
      const int d = idx & 3, t = (idx >> 2) & 3, kt = idx >> 4;
      const int k = 4 * kt + t, f = k / 15, j = k - 15 * f, i = 32 + d;
      double v = 0.0;
      if (d < kNL) v = FE ? Og[(i * 4 + f) * 15 + j] : Og[(f * 35 + i) * 15 + j];
      sL[idx] = v;
    }
  }
  __syncthreads();
  FNSM_TL(2);
  stagger_start(warp, flags >> 8);
  for (uint32_t n = 0; cur < nchunks; ++n) {
    mbar_wait(bar, n & 1u);
    unsigned tk = 0;
    if (FS) tk = wq.ticket(lane);
    if (n == 0) FNSM_TL(3);
    double a[kME][L::KT];
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      const int el = chunk_el(g, m);
      double Jf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) Jf[f] = FE ? sJ[f * kCH + el] : sJ[el * 4 + f];
      const double* sv = s + el * 15 + t;
      int sh4[4];                                // bulk instantiation: where face slab f starts inside its region
#pragma unroll
      for (int f = 0; f < 4; ++f) sh4[f] = TMA ? 0 : (shm >> f) & 1;
#pragma unroll
      for (int kt = 0; kt < L::KT; ++kt) {
        // k = 4kt + t = 15 f + j: f = F0 for t < THR, F0 + 1 above (THR >= 4: the tile does not straddle a face)
        const int F0 = (4 * kt) / 15, THR = 15 * (F0 + 1) - 4 * kt;
        // address of v[f][el][j] = f*V_SLAB + el*15 + (k - 15 f) = el*15 + t + 4kt + (V_SLAB - 15) f
        const int off0 = 4 * kt + (P - 15) * F0;
        double jf = Jf[F0], v;
        if (THR < 4) {
          const bool up = t >= THR;
          if constexpr (TMA) v = sv[off0 + (up ? (L::V_SLAB - 15) : 0)];
          else               v = sv[off0 + (up ? (P - 15) + sh4[F0 + 1 < 4 ? F0 + 1 : 3] : sh4[F0])];
          jf = up ? Jf[F0 + 1 < 4 ? F0 + 1 : 3] : jf;
        } else {
          if constexpr (TMA) v = sv[off0];
          else               v = sv[off0 + sh4[F0]];
        }
        a[m][kt] = jf * v;
      }
    }
    __syncwarp();
    if (FS) nxt = wq.resolve(tk);
    // next item: (FS) the next queue position, else the next field of this chunk / field 0 of the next chunk
    bool advance;
    int nfld;
    long long nchunk;
    if (FS) {
      advance = true;
      const unsigned q = nrows > 1 ? __umulhi((unsigned)nxt, magic) : (unsigned)nxt;
      nfld = (int)((unsigned)nxt - q * (unsigned)nrows);
      nchunk = q;
    } else {
      advance = fld + 1 == nrows;
      nfld = advance ? 0 : fld + 1;
      nchunk = advance ? nxt : cur;
    }
    int shn = 0;
    if (nchunk < nchunks) {
      const double* vg = static_cast<const double*>(rows.field[nfld]);
      if constexpr (TMA) {
        lift_issue<FE>(s, bar, &maps.in[nfld], &maps.jac, Jg, vg, nchunk, E, tma, lane);
      } else {
        shn = lift_shift_mask(vg, E);
        lift_issue_bulk<FE>(s, bar, Jg, vg, nchunk, nchunks, E, lane, shn);
      }
    }

    if (!FS && advance) tk = wq.ticket(lane);      // ticket after next; its latency hides under the DMMAs

    double acc[kME][kNT][2];
#ifndef FNSM_LIFT_LP
#define FNSM_LIFT_LP 2
#endif
    constexpr int kLLP = FNSM_LIFT_LP;
    double accL[kLLP][kME][kNL];   // partial sums per value: DFMA latency >> 6 chains
#pragma unroll
    for (int m = 0; m < kME; ++m) {
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) { acc[m][nt][0] = 0.0; acc[m][nt][1] = 0.0; }
#pragma unroll
      for (int d = 0; d < kNL; ++d)
#pragma unroll
        for (int p = 0; p < kLLP; ++p) accL[p][m][d] = 0.0;
    }
    {
#pragma unroll
    for (int kt = 0; kt < L::KT; ++kt) {
      if constexpr (!TMA) {                      // paired fragments, see lift_b_pos
        const double2 b01 = *reinterpret_cast<const double2*>(sB + ((kt * 2 + 0) * 32 + lane) * 2);
        const double2 b23 = *reinterpret_cast<const double2*>(sB + ((kt * 2 + 1) * 32 + lane) * 2);
#pragma unroll
        for (int m = 0; m < kME; ++m) {
          dmma884(acc[m][0], a[m][kt], b01.x);
          dmma884(acc[m][1], a[m][kt], b01.y);
          dmma884(acc[m][2], a[m][kt], b23.x);
          dmma884(acc[m][3], a[m][kt], b23.y);
        }
      } else {
        const double* bp = sB + (kt * kNT) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < kNT; ++nt) {
          const double b = bp[nt * 32];
#pragma unroll
          for (int m = 0; m < kME; ++m) dmma884(acc[m][nt], a[m][kt], b);
        }
      }
      const double2 l01 = *reinterpret_cast<const double2*>(sL + (kt * 4 + t) * 4);
      const double l2 = sL[(kt * 4 + t) * 4 + 2];
#pragma unroll
      for (int m = 0; m < kME; ++m) {
        accL[kt % kLLP][m][0] = fma(a[m][kt], l01.x, accL[kt % kLLP][m][0]);
        accL[kt % kLLP][m][1] = fma(a[m][kt], l01.y, accL[kt % kLLP][m][1]);
        accL[kt % kLLP][m][2] = fma(a[m][kt], l2, accL[kt % kLLP][m][2]);
      }
    }
    }

    const long long c = cur;
    const long long e0 = c * kCH;
    double* outg = static_cast<double*>(rows.out[fld]);
    const int so = TMA ? 0 : odd8(outg);           // bulk stores: the block is staged from stage + so on
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int m = 0; m < kME; ++m) {
      double* o = stage + so + chunk_el(g, m) * 35;
      double ls[kNL];
#pragma unroll
      for (int d = 0; d < kNL; ++d) {
        ls[d] = accL[0][m][d];
#pragma unroll
        for (int p = 1; p < kLLP; ++p) ls[d] += accL[p][m][d];
      }
      const double l0 = quad_sum(ls[0]), l1 = quad_sum(ls[1]), l2 = quad_sum(ls[2]);
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) {
        o[8 * nt + 2 * t] = acc[m][nt][0];
        o[8 * nt + 2 * t + 1] = acc[m][nt][1];
      }
      if (t < kNL) o[32 + t] = t == 0 ? l0 : (t == 1 ? l1 : l2);
    }
    fence_proxy_async();
    __syncwarp();
    if constexpr (TMA) {
      if (tma) {
        if (lane == 0) { tma_store_2d(&maps.out[fld], stage, 0, (int)(c * (kCH / 2))); tma_store_commit(); }
      } else {
        flush_plain(outg + e0 * 35, stage, e0, E, lane);
      }
    } else if (c + 1 < nchunks) {
      // whole block: one bulk store; a run that starts 8 (mod 16) leaves without its first and last double, which
      // two lanes store directly (see k_grad_dmma)
      if (lane == 0) {
        bulk_store(outg + e0 * 35 + so, stage + 2 * so, OUT_BLOCK * 8 - 16 * so);
        tma_store_commit();
      }
      if (so && lane >= 1 && lane < 3) {
        const int k = lane == 1 ? 0 : OUT_BLOCK - 1;
        outg[e0 * 35 + k] = stage[1 + k];
      }
    } else {
      flush_plain(outg + e0 * 35, stage + so, e0, E, lane);
    }
    if (n == 0) FNSM_TL(4);
    FNSM_TL_VAL(7, n + 1);
    if (FS) cur = nchunk;
    else if (advance) { cur = nchunk; nxt = wq.resolve(tk); }
    fld = nfld;
    shm = shn;
  }
  FNSM_TL(5);
  if (lane == 0) tma_store_wait_all();
  FNSM_TL(6);
  wait_for_previous_kernels();   // kFlagIndependent: the only wait; otherwise a no-op (the grid ahead has completed)
}

// ------------------------------------------------------------ launchers ----
inline bool dmma_supported(int kind, int n_outer, int ni, int nj) {
  if (kind == FNSM_OP_GRAD || kind == FNSM_OP_DIV) return n_outer == 3 && ni == 35 && nj == 35;
  return n_outer == 4 && ni == 35 && nj == 15;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// opt in to > 48 KB of dynamic shared memory; remembered per (kernel address, device) so the steady-state
// launch path makes no attribute call (the cap is only ever raised).  Keyed by the function ADDRESS: every
// instantiation of one kernel template has the same pointer type, so a per-type static would be shared.
struct SmemGrant { std::atomic<const void*> tag; std::atomic<size_t> bytes; };
template <class K>
static int set_smem(K kernel, size_t smem) {
  constexpr int kSlots = 256;
  static SmemGrant grants[kSlots];
  int dev = 0;
  cudaGetDevice(&dev);
  const void* fn = reinterpret_cast<const void*>(kernel);
  const uint64_t h = ((uint64_t)reinterpret_cast<uintptr_t>(fn) >> 4) * 0x9E3779B97F4A7C15ull + (uint64_t)dev * 0x632BE5ABull;
  SmemGrant& g = grants[(h >> 32) % kSlots];
  // the slot is tagged with fn + dev; a colliding kernel simply re-issues the attribute call
  const void* tag = static_cast<const char*>(fn) + dev;
  if (g.tag.load(std::memory_order_acquire) == tag && g.bytes.load(std::memory_order_relaxed) >= smem) return FNSM_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  g.tag.store(nullptr, std::memory_order_release);
  g.bytes.store(smem, std::memory_order_relaxed);
  g.tag.store(tag, std::memory_order_release);
  return FNSM_OK;
}

// cuTensorMapEncodeTiled, fetched through the runtime (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// fp64 tensor map of rank 2 or 3; dims / box innermost first, strides (bytes) of dims 1..rank-1
// Encoding a descriptor costs ~1 us of host time and a launch needs 3-9 of them; timing loops and
// time-stepping codes call with the same buffers over and over, so the last encodings are kept in a
// small per-thread cache keyed by everything the encoder sees.
struct MapKey {
  const void* base; int dtype, rank; cuuint64_t dims[3], strides[2]; cuuint32_t box[3];
  bool operator==(const MapKey& o) const {
    if (base != o.base || dtype != o.dtype || rank != o.rank) return false;
    for (int k = 0; k < rank; ++k) if (dims[k] != o.dims[k] || box[k] != o.box[k]) return false;
    for (int k = 0; k + 1 < rank; ++k) if (strides[k] != o.strides[k]) return false;
    return true;
  }
};
static bool make_map_typed(CUtensorMap* tm, CUtensorMapDataType dtype, const void* base, int rank,
                           const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  constexpr int N = 32;
  thread_local MapKey keys[N];
  thread_local CUtensorMap maps[N];
  thread_local int used = 0, next = 0;
  MapKey key{};
  key.base = base; key.dtype = (int)dtype; key.rank = rank;
  for (int k = 0; k < rank; ++k) { key.dims[k] = dims[k]; key.box[k] = box[k]; }
  for (int k = 0; k + 1 < rank; ++k) key.strides[k] = strides[k];
  for (int i = 0; i < used; ++i)
    if (keys[i] == key) { *tm = maps[i]; return true; }
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint32_t ones[3] = {1, 1, 1};
  if (enc(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, ones,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  keys[next] = key; maps[next] = *tm;
  next = (next + 1) % N;
  if (used < N) ++used;
  return true;
}
static bool make_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                     const cuuint64_t* strides, const cuuint32_t* box) {
  return make_map_typed(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, base, rank, dims, strides, box);
}
// (E, W) row-major with W*2 doubles per element pair: box = 8 pairs = one 16-element chunk
static bool map_rows(CUtensorMap* tm, const void* base, long long E, int W) {
  const cuuint64_t dims[2] = {(cuuint64_t)(2 * W), (cuuint64_t)(E / 2)};
  const cuuint64_t strides[1] = {(cuuint64_t)(16 * W)};
  const cuuint32_t box[2] = {(cuuint32_t)(2 * W), (cuuint32_t)(kCH / 2)};
  return make_map(tm, base, 2, dims, strides, box);
}
// (S, E, W): S slabs of (E, W)
static bool map_slabs(CUtensorMap* tm, const void* base, long long E, int W, int S) {
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * W), (cuuint64_t)(E / 2), (cuuint64_t)S};
  const cuuint64_t strides[2] = {(cuuint64_t)(16 * W), (cuuint64_t)E * W * 8};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * W), (cuuint32_t)(kCH / 2), (cuuint32_t)S};
  return make_map(tm, base, 3, dims, strides, box);
}
// (R, E): R rows with the element axis contiguous
static bool map_erows(CUtensorMap* tm, const void* base, long long E, int R) {
  const cuuint64_t dims[2] = {(cuuint64_t)E, (cuuint64_t)R};
  const cuuint64_t strides[1] = {(cuuint64_t)E * 8};
  const cuuint32_t box[2] = {(cuuint32_t)kCH, (cuuint32_t)R};
  return make_map(tm, base, 2, dims, strides, box);
}

// Kernel launch through cudaLaunchKernelEx.  Inside fnsm_b200_wave3d_fused the second and third kernel are
// launched with programmatic stream serialization: the three einsums of the wave operator neither read nor
// write each other's buffers, every kernel releases its dependents at its first instruction
// (griddepcontrol.launch_dependents), so the next kernel's CTAs move onto an SM as soon as the previous
// kernel's persistent CTA there exits -- its tail overlaps the next prologue (operator tables, first TMA
// loads).  The first kernel of the operator and everything launched after it keep normal stream order.
inline bool& overlap_with_previous_kernel() {
  thread_local bool flag = false;
  return flag;
}
// `pdl_always` (the three kernels of this file): every launch carries the attribute.  These kernels execute
// griddepcontrol.wait before their first global access unless kFlagIndependent is set, so the result is ordinary
// stream order with the launch latency and the on-chip set-up hidden behind the previous kernel's tail (3.6 us
// between two launches without it, tools/timeline) -- whatever the previous kernel is: a grid that never releases
// its dependents releases them when it completes.
inline int independent_flag() { return overlap_with_previous_kernel() ? kFlagIndependent : 0; }
template <bool pdl_always = false, class... KArgs, class... Args>
static void launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t st,
                     Args&&... args) {
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(threads);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_always || overlap_with_previous_kernel()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
  }
  cudaLaunchKernelEx(&lc, kernel, std::forward<Args>(args)...);   // errors are picked up by post_launch()
}

// "Fast start" instantiations (FS = true, see k_div_dmma) for launches that leave a warp only a few work items:
// there the 4-5 us they save per launch outweigh their slightly slower steady state.  fs_mode 1 / 2
// force them on / off (cfg->reserved[2] bits 4 / 5: A/B runs and tests).  Compiled for the default warp counts only.
// Crossovers measured on B200 (profiles/r02_small_e.md): grad ~30 chunks per warp (E ~ 700 k), div ~20 (E ~ 570 k),
// lift ~8 (E ~ 300 k; its fine queue costs more per item).
inline bool fast_start(int kind, long long nchunks, int sms, int nw, int fs_mode) {
  if (fs_mode) return fs_mode == 1;
  const int below = kind == FNSM_OP_GRAD ? 30 : (kind == FNSM_OP_DIV ? 20 : 8);
  return nchunks < (long long)below * sms * nw;
}
inline int fs_mode_of(const fnsm_cfg* cfg) { return cfg ? ((cfg->reserved[2] >> 4) & 3) : 0; }

// Operands that do not qualify for tensor maps (odd E, a base that is not 16-byte aligned): the TMA = false
// instantiations, at the default warp counts (div 12 with direct stores, grad 10, lift 16).
static int launch_dmma_plain(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                             long long E, int stagger, const DevInfo& di, cudaStream_t st, int fs_mode = 0) {
  const long long nchunks = (E + kCH - 1) / kCH;
  const double* J = static_cast<const double*>(jac);
  const double* O = static_cast<const double*>(op);
  auto grid_for = [&](int nw, long long nitems) {
    const long long need = (nitems + nw - 1) / nw;
    return (unsigned)(di.sms < need ? di.sms : need);
  };
  const int flags = (stagger << 8) | independent_flag();
  OpMaps maps{};
  if (kind == FNSM_OP_DIV || kind == FNSM_OP_GRAD) {
    const bool is_div = kind == FNSM_OP_DIV;
    const int NW = is_div ? 12 : 10;
    const size_t smem = is_div ? 8 * ((size_t)DivLayout::B_DOUBLES + (size_t)NW * DivLayout::SLOT_DOUBLES_BULK) + 8 * (size_t)NW + 8
                               : 8 * ((size_t)GradLayout::B_DOUBLES + (size_t)NW * (GradLayout::SLOT_DOUBLES + 2 + 3 * (OUT_BLOCK + 2))) + 8 * (size_t)NW + 8;
    const bool fs = fast_start(kind, nchunks, di.sms, NW, fs_mode);
    auto go = [&](auto kernel, auto pdl) {
      if (int rc = set_smem(kernel, smem)) return rc;
      for (int r = 0; r < nrows; ++r) {
        launch_k<decltype(pdl)::value>(kernel, grid_for(NW, nchunks), NW * 32, smem, st, maps, J, O,
                                       static_cast<const double*>(rows.field[r]), static_cast<double*>(rows.out[r]), E, flags);
        if (int rc = post_launch()) return rc;
      }
      return (int)FNSM_OK;
    };
    if (is_div) return fs ? go(k_div_dmma<12, false, 0, 3, false, false, true>, std::true_type{})
                          : go(k_div_dmma<12, false, 0, 3, false, false, false>, std::false_type{});
    return fs ? go(k_grad_dmma<10, false, true>, std::true_type{}) : go(k_grad_dmma<10, false, false>, std::false_type{});
  }
  constexpr int NW = 16;
  const size_t smem = 8 * ((size_t)LiftLayout::B_DOUBLES + (size_t)NW * (LiftLayout::SLOT_DOUBLES + 8 + OUT_BLOCK + 2)) + 8 * (size_t)NW + 8;
  LiftMaps lmaps{};
  const bool fs = fast_start(kind, nchunks, di.sms, NW, fs_mode);
  auto go = [&](auto kernel, auto pdl) {
    if (int rc = set_smem(kernel, smem)) return rc;
    launch_k<decltype(pdl)::value>(kernel, grid_for(NW, fs ? nchunks * nrows : nchunks), NW * 32, smem, st, lmaps, J, O,
                                   rows, nrows, E, flags);
    return post_launch();
  };
  const bool fe = kind == FNSM_OP_LIFT_FE;
  if (fs) return fe ? go(k_lift_dmma<NW, true, false, true>, std::true_type{}) : go(k_lift_dmma<NW, false, false, true>, std::true_type{});
  return fe ? go(k_lift_dmma<NW, true, false, false>, std::false_type{}) : go(k_lift_dmma<NW, false, false, false>, std::false_type{});
}

template <int NW>
static int launch_dmma_nw(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                          long long E, const fnsm_cfg* cfg, const DevInfo& di, cudaStream_t st) {
  const long long nchunks = (E + kCH - 1) / kCH;
  const double* J = static_cast<const double*>(jac);
  const double* O = static_cast<const double*>(op);
  constexpr int threads = NW * 32;
  // TMA path: 16-B aligned bases and rows (E even); int32 box coordinates
  bool tma = (E % 2 == 0) && E < (1LL << 31) - kCH && aligned16(jac);
  for (int r = 0; r < nrows; ++r) tma = tma && aligned16(rows.field[r]) && aligned16(rows.out[r]);
  const int dbg = cfg ? cfg->reserved[0] : 0;       // bit 0: force the plain path; bits 1, 2: profiling aids
  if (dbg & 1) tma = false;
  int stagger = cfg ? cfg->reserved[1] : 0;           // start-up phase offset (cycles); 0 = library default
  if (stagger == 0) stagger = kDefaultStagger;
  if (stagger < 0) stagger = 0;                       // negative: off
  if (stagger > (1 << 20)) stagger = 1 << 20;
  const int fs_mode = fs_mode_of(cfg);
  if (!tma) return launch_dmma_plain(kind, jac, op, rows, nrows, E, stagger, di, st, fs_mode);
  auto grid_for_items = [&](long long nitems) {
    long long grid = di.sms;                          // one persistent CTA per SM
    const long long need = (nitems + NW - 1) / NW;      // no more CTAs than can be kept busy
    return (unsigned)(grid < need ? grid : need);
  };
  const bool fs = fast_start(kind, nchunks, di.sms, NW, fs_mode);

  if (kind == FNSM_OP_DIV || kind == FNSM_OP_GRAD) {
    const bool is_div = kind == FNSM_OP_DIV;
    const size_t slot_d = is_div ? DivLayout::SLOT_DOUBLES : GradLayout::SLOT_DOUBLES;
    const size_t b_d = is_div ? DivLayout::B_DOUBLES : GradLayout::B_DOUBLES;
    size_t stage_d = is_div ? OUT_BLOCK : 3 * OUT_BLOCK;
    size_t smem = 8 * (b_d + (size_t)NW * (slot_d + stage_d)) + 8 * (size_t)NW + 8;
    bool staged = true;
    if (is_div && (smem > (size_t)di.max_smem_optin || (cfg && (cfg->reserved[0] & 8)))) {
      staged = false;                                // no room for the output stage: direct stores
      stage_d = 0;
      smem = 8 * (b_d + (size_t)NW * slot_d) + 8 * (size_t)NW + 8;
    }
    if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
    for (int r = 0; r < nrows; ++r) {
      const double* u = static_cast<const double*>(rows.field[r]);
      double* out = static_cast<double*>(rows.out[r]);
      OpMaps maps;
      bool ok = tma && map_erows(&maps.jac, J, E, 9);
      if (is_div) ok = ok && map_slabs(&maps.in, u, E, 35, 3) && map_rows(&maps.out, out, E, 35);
      else        ok = ok && map_rows(&maps.in, u, E, 35) && map_slabs(&maps.out, out, E, 35, 3);
      if (!ok) {                                        // the driver refused a descriptor: plain producer for this row
        OpmatRows one{};
        one.field[0] = u; one.out[0] = out;
        if (int rc = launch_dmma_plain(kind, jac, op, one, 1, E, stagger, di, st, fs_mode)) return rc;
        continue;
      }
      const int flags = kFlagTma | (dbg & (kFlagNoLoad | kFlagNoStore)) | (stagger << 8) | independent_flag();
      auto go = [&](auto kernel, auto pdl) {
        if (int rc = set_smem(kernel, smem)) return rc;
        launch_k<decltype(pdl)::value>(kernel, grid_for_items(nchunks), threads, smem, st, maps, J, O, u, out, E, flags);
        return post_launch();
      };
      int rc;
      if (is_div) {
        const int dbgk = cfg ? (cfg->reserved[2] & 15) : 0;
        if (staged && NW == 10 && dbgk) {
#define FNSM_DBG_CASE(M) case M: rc = go(k_div_dmma<10, true, M>, std::false_type{}); break;
          switch (dbgk) { FNSM_DBG_CASE(1) FNSM_DBG_CASE(2) FNSM_DBG_CASE(4) FNSM_DBG_CASE(3) FNSM_DBG_CASE(5) FNSM_DBG_CASE(6) FNSM_DBG_CASE(7) default: return FNSM_E_BAD_CONFIG; }
#undef FNSM_DBG_CASE
        } else if (staged) {
          rc = go(k_div_dmma<NW, true>, std::false_type{});
        } else {
          if constexpr (NW == 12) {
            rc = fs ? go(k_div_dmma<NW, false, 0, 3, false, true, true>, std::true_type{}) : go(k_div_dmma<NW, false>, std::false_type{});
          } else {
            rc = go(k_div_dmma<NW, false>, std::false_type{});
          }
        }
      } else {
        if constexpr (NW == 10) {
          rc = fs ? go(k_grad_dmma<NW, true, true>, std::true_type{}) : go(k_grad_dmma<NW>, std::false_type{});
        } else {
          rc = go(k_grad_dmma<NW>, std::false_type{});
        }
      }
      if (rc) return rc;
    }
    return FNSM_OK;
  }
  const size_t smem = 8 * ((size_t)LiftLayout::B_DOUBLES + (size_t)NW * (LiftLayout::SLOT_DOUBLES + OUT_BLOCK)) + 8 * (size_t)NW + 8;
  if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  const bool fine = NW == 16 && fs;                   // compiled for the default warp count
  const unsigned grid = grid_for_items(fine ? nchunks * nrows : nchunks);
  LiftMaps maps;
  bool ok = tma && (kind == FNSM_OP_LIFT_FE ? map_erows(&maps.jac, J, E, 4) : map_rows(&maps.jac, J, E, 4));
  for (int r = 0; r < nrows && ok; ++r)
    ok = map_slabs(&maps.in[r], rows.field[r], E, 15, 4) && map_rows(&maps.out[r], rows.out[r], E, 35);
  if (!ok) return launch_dmma_plain(kind, jac, op, rows, nrows, E, stagger, di, st, fs_mode);
  const int flags = kFlagTma | (stagger << 8) | independent_flag();
  auto go = [&](auto kernel, auto pdl) {
    if (int rc = set_smem(kernel, smem)) return rc;
    launch_k<decltype(pdl)::value>(kernel, grid, threads, smem, st, maps, J, O, rows, nrows, E, flags);
    return post_launch();
  };
  const bool fe = kind == FNSM_OP_LIFT_FE;
  if constexpr (NW == 16) {
    if (fine) return fe ? go(k_lift_dmma<NW, true, true, true>, std::true_type{}) : go(k_lift_dmma<NW, false, true, true>, std::true_type{});
  }
  return fe ? go(k_lift_dmma<NW, true>, std::false_type{}) : go(k_lift_dmma<NW, false>, std::false_type{});
}

static int launch_dmma(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                       int n_outer, int ni, int nj, long long E, const fnsm_cfg* cfg,
                       const DevInfo& di, cudaStream_t st) {
  (void)n_outer; (void)ni; (void)nj;
  if (cfg && cfg->ctas_per_sm > 1) return FNSM_E_BAD_CONFIG;
  if (cfg && (cfg->stages < 0 || cfg->stages > 2)) return FNSM_E_BAD_CONFIG;
  // defaults from the round-1 sweeps on B200 (profiles/): grad 10 warps, div 12 (direct stores), lift 16
  // (lift fits 16 warps in shared memory and, at 126 registers, in the register file: 79.6 % -> 81.1 %)
  const int dflt = kind == FNSM_OP_GRAD ? 320 : (kind == FNSM_OP_DIV ? 384 : 512);
  const int threads = (cfg && cfg->threads != 0) ? cfg->threads : dflt;
  switch (threads) {
    case 128: return launch_dmma_nw<4>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 256: return launch_dmma_nw<8>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 288: return launch_dmma_nw<9>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 320: return launch_dmma_nw<10>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 352: return launch_dmma_nw<11>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 384: return launch_dmma_nw<12>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 448: return launch_dmma_nw<14>(kind, jac, op, rows, nrows, E, cfg, di, st);
    case 512: return launch_dmma_nw<16>(kind, jac, op, rows, nrows, E, cfg, di, st);
    default: return FNSM_E_BAD_CONFIG;
  }
}

// wave_3d_p4: the three kernels back to back on the caller's stream -- the structure of the reference, whose
// single loopy translation unit is split into three device kernels (examples/wave_3d_p4_auto.py:36-56).  A
// single persistent kernel was sized and rejected (shared memory for 6 instead of 10-12
// warps per SM; the einsums share only J, 72 of 5 456 B per element).
static int launch_wave3d_dmma(const fnsm_wave_args* a, long long E, const fnsm_cfg* cfg,
                              const DevInfo& di, cudaStream_t st) {
  OpmatRows r1{}; r1.field[0] = a->v; r1.out[0] = a->div_out;
  if (int rc = launch_dmma(FNSM_OP_DIV, a->J, a->D, r1, 1, 3, 35, 35, E, cfg, di, st)) return rc;
  OpmatRows r2{}; r2.field[0] = a->u; r2.out[0] = a->grad_out;
  OpmatRows r3{};
  for (int k = 0; k < 4; ++k) { r3.field[k] = a->F[k]; r3.out[k] = a->lift_out[k]; }
  overlap_with_previous_kernel() = true;                 // independent einsums: see launch_k
  int rc = launch_dmma(FNSM_OP_GRAD, a->J, a->D, r2, 1, 3, 35, 35, E, cfg, di, st);
  if (!rc) rc = launch_dmma(FNSM_OP_LIFT_FE, a->Jface, a->L, r3, 4, 4, 35, 15, E, cfg, di, st);
  overlap_with_previous_kernel() = false;
  return rc;
}

}  // namespace fnsm
