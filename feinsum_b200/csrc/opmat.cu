// C-ABI entry points of the operator-matrix x element-batch einsum classes
// (DG gradient / divergence / face-mass lift) and of the fused wave_3d_p4
// operator; dispatch between kernel variants.
//
// replaces: reference generate_loopy + tuning/impls/{xre_rij_ej_to_xei,
// xre_rij_xej_to_ei*, ifj_fe_fej_to_ei*, batched_*}.py + the executor launch
// in src/feinsum/measure.py:244-273.
#include "common.cuh"
#include "opmat_simt.cuh"
#include "opmat_dmma.cuh"
#include "opmat_grad2.cuh"
#include "opmat_tf32.cuh"
#include "opmat_tc32.cuh"
#include "opmat_dmma_gen.cuh"
#include "opmat_tf32_gen.cuh"
#include "opmat_se.cuh"
#include <cstdio>
#include <cstring>

namespace fnsm {

static void set_range(fnsm_cfg_range* r, const char* name, int lo, int hi, int step, int dflt) {
  std::memset(r, 0, sizeof(*r));
  std::snprintf(r->name, sizeof(r->name), "%s", name);
  r->lo = lo; r->hi = hi; r->step = step; r->dflt = dflt;
}

int opmat_cfg_space(int kernel_id, fnsm_cfg_range* out, int cap) {
  (void)kernel_id;
  fnsm_cfg_range tmp[5];
  int n = 0;
  // 0 = auto, 1 = mma.sync tensor path (DMMA / 3xTF32, p=4 shapes), 2 = simt, 3 = tcgen05 3xTF32 (fp32, p=4 shapes)
  set_range(&tmp[n++], "variant", 0, 3, 1, 0);
  set_range(&tmp[n++], "tile_e", 8, 64, 8, 16);        // simt: elements per CTA tile
  set_range(&tmp[n++], "ctas_per_sm", 1, 8, 1, 0);     // persistent grid size
  set_range(&tmp[n++], "threads", 128, 512, 32, 0);   // dmma: 32 * warps per persistent CTA (4, 8..12, 14, 16)
  set_range(&tmp[n++], "stages", 0, 2, 1, 1);          // dmma: shared-memory slots per warp; 2 = grad2 formulation (fp64 grad p = 4)
  for (int i = 0; i < n && i < cap; ++i) out[i] = tmp[i];
  return n;
}

template <typename T>
static int launch_simt(int kind, const void* jac, const void* op, const OpmatRows& rows, int nrows,
                       int n_outer, int ni, int nj, long long E, const fnsm_cfg* cfg,
                       const DevInfo& di, cudaStream_t st) {
  // default tile: at least two (element, dof) work items per thread of the 256-thread CTA
  int tile_e = (cfg && cfg->tile_e > 0) ? cfg->tile_e : 16;
  if (!(cfg && cfg->tile_e > 0)) {
    const int want = (512 + ni - 1) / ni;
    tile_e = want <= 16 ? 16 : (want >= 128 ? 128 : (want + 7) / 8 * 8);
  }
  if (tile_e < 1 || tile_e > 256) return FNSM_E_BAD_CONFIG;
  if (n_outer > 4 && (kind == FNSM_OP_GRAD || kind == FNSM_OP_DIV)) return FNSM_E_UNSUPPORTED;
  size_t smem = 0;
  if (kind == FNSM_OP_GRAD)
    smem = sizeof(T) * ((size_t)n_outer * ni * nj + (size_t)tile_e * nj + (size_t)n_outer * n_outer * tile_e);
  else if (kind == FNSM_OP_DIV)
    smem = sizeof(T) * ((size_t)n_outer * ni * nj + (size_t)n_outer * tile_e * nj + (size_t)n_outer * n_outer * tile_e);
  else
    smem = sizeof(T) * ((size_t)n_outer * ni * nj + (size_t)n_outer * tile_e * nj + (size_t)n_outer * tile_e);
  if (smem > (size_t)di.max_smem_optin) return FNSM_E_BAD_CONFIG;
  const long long ntiles = (E + tile_e - 1) / tile_e;
  int cps = (cfg && cfg->ctas_per_sm > 0) ? cfg->ctas_per_sm : 4;
  long long grid = (long long)cps * di.sms;
  if (grid > ntiles) grid = ntiles;
  const T* J = static_cast<const T*>(jac);
  const T* O = static_cast<const T*>(op);
  cudaError_t e = cudaSuccess;
#define FNSM_LAUNCH(KERNEL)                                                                  \
  do {                                                                                       \
    e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return (int)e;                                                     \
    KERNEL<<<(unsigned)grid, 256, smem, st>>>(J, O, rows, nrows, n_outer, ni, nj, E, tile_e); \
  } while (0)
  switch (kind) {
    case FNSM_OP_GRAD: FNSM_LAUNCH(k_grad_simt<T>); break;
    case FNSM_OP_DIV: FNSM_LAUNCH(k_div_simt<T>); break;
    case FNSM_OP_LIFT_EF: FNSM_LAUNCH((k_lift_simt<T, false>)); break;
    case FNSM_OP_LIFT_FE: FNSM_LAUNCH((k_lift_simt<T, true>)); break;
    default: return FNSM_E_BAD_ARG;
  }
#undef FNSM_LAUNCH
  return post_launch();
}

static int opmat_dispatch(int kind, int dtype, const void* jac, const void* op,
                          const void* const* fields, void* const* outs, int b,
                          int n_outer, int ni, int nj, long long E, const fnsm_cfg* cfg,
                          cudaStream_t st) {
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  int variant = cfg ? cfg->variant : 0;    // 0: library default
  if (variant < 0 || variant > 3) return FNSM_E_BAD_CONFIG;
  // variant 1 = mma.sync tensor path (fp64 DMMA, fp32 3xTF32; p = 4 tuned, p = 1..3 generic), 2 = simt (any shape),
  // 3 = tcgen05 3xTF32 (fp32, tets p = 1..4; TMA producer when E % 4 == 0 and the bases are 16-byte aligned,
  //     cp.async producer + plain vector stores otherwise)
  const bool tensor_ok = dmma_supported(kind, n_outer, ni, nj);
  const bool tc_ok = dtype == FNSM_F32 && tc32_supported(kind, n_outer, ni, nj);   // tets p = 1..4
  const bool gen_ok = dmma_gen_supported(kind, n_outer, ni, nj);   // generic warp-per-chunk tensor kernels, tets p = 1..3
  if (variant == 1 && !tensor_ok && !gen_ok) return FNSM_E_UNSUPPORTED;
  if (variant == 3 && !tc_ok) return FNSM_E_UNSUPPORTED;
  const bool is_auto = variant == 0;
  // auto: fp32 tets p = 1..4 -> tcgen05;
  // fp64 -> DMMA (tuned p = 4 kernels, generic p = 1..3 kernel); every other shape -> simt
  if (is_auto) variant = tc_ok ? 3 : ((tensor_ok || gen_ok) ? 1 : 2);
  for (int r0 = 0; r0 < b; r0 += 8) {
    const int nr = (b - r0 < 8) ? (b - r0) : 8;
    OpmatRows rows{};
    for (int r = 0; r < nr; ++r) {
      rows.field[r] = fields[r0 + r];
      rows.out[r] = outs[r0 + r];
      if (!rows.field[r] || !rows.out[r]) return FNSM_E_BAD_ARG;
    }
    int rc;
    if (variant == 1 && dtype == FNSM_F64 && tensor_ok && kind == FNSM_OP_GRAD && cfg && cfg->stages == 2) {
      // grad2 (opmat_grad2.cuh): divergence tables + direct stores; selected with stages = 2 while it is being measured
      const int th = cfg->threads ? cfg->threads : 384;
      const bool fp = cfg->reserved[0] & 1;
      const double* Jd = static_cast<const double*>(jac);
      const double* Dd = static_cast<const double*>(op);
      rc = th == 384 ? launch_grad2<12>(Jd, Dd, rows, nr, E, fp, di, st)
         : th == 320 ? launch_grad2<10>(Jd, Dd, rows, nr, E, fp, di, st)
         : th == 512 ? launch_grad2<16>(Jd, Dd, rows, nr, E, fp, di, st)
         : th == 448 ? launch_grad2<14>(Jd, Dd, rows, nr, E, fp, di, st) : (int)FNSM_E_BAD_CONFIG;
    } else if (variant == 1 && dtype == FNSM_F64)
      rc = tensor_ok ? launch_dmma(kind, jac, op, rows, nr, n_outer, ni, nj, E, cfg, di, st)
                     : launch_dmma_gen(kind, jac, op, rows, nr, ni, E, cfg, di, st);
    else if (variant == 3)
      rc = launch_tc32(kind, jac, op, rows, nr, ni, E, di, st, cfg && (cfg->reserved[0] & 1));
    else if (variant == 1)
      rc = tensor_ok ? launch_tf32(kind, jac, op, rows, nr, E, cfg, di, st)
                     : launch_tf32_gen(kind, jac, op, rows, nr, ni, E, cfg, di, st);
    else if (dtype == FNSM_F64)
      rc = launch_simt<double>(kind, jac, op, rows, nr, n_outer, ni, nj, E, cfg, di, st);
    else
      rc = launch_simt<float>(kind, jac, op, rows, nr, n_outer, ni, nj, E, cfg, di, st);
    if (rc) return rc;
  }
  return FNSM_OK;
}

}  // namespace fnsm

extern "C" int fnsm_b200_opmat_batch(int32_t kind, int32_t dtype, const void* jac, const void* op,
                                     const void* const* fields, void* const* outs, int32_t b,
                                     int32_t n_outer, int32_t n_i, int32_t n_j, int64_t E,
                                     const fnsm_cfg* cfg, void* stream) {
  using namespace fnsm;
  if (!jac || !op || !fields || !outs || b <= 0 || E < 0) return FNSM_E_BAD_ARG;
  if (kind < FNSM_OP_GRAD || kind > FNSM_OP_LIFT_FE) return FNSM_E_BAD_ARG;
  if (dtype != FNSM_F64 && dtype != FNSM_F32) return FNSM_E_UNSUPPORTED;
  if (n_outer < 1 || n_i < 1 || n_j < 1) return FNSM_E_BAD_ARG;
  if (E == 0) return FNSM_OK;
  return opmat_dispatch(kind, dtype, jac, op, fields, outs, b, n_outer, n_i, n_j, E, cfg,
                        static_cast<cudaStream_t>(stream));
}

extern "C" int fnsm_b200_wave3d_fused(int32_t dtype, const fnsm_wave_args* a, int64_t E,
                                      const fnsm_cfg* cfg, void* stream) {
  using namespace fnsm;
  if (!a || E < 0) return FNSM_E_BAD_ARG;
  if (dtype != FNSM_F64 && dtype != FNSM_F32) return FNSM_E_UNSUPPORTED;
  if (!a->J || !a->D || !a->v || !a->u || !a->L || !a->Jface || !a->div_out || !a->grad_out)
    return FNSM_E_BAD_ARG;
  for (int k = 0; k < 4; ++k)
    if (!a->F[k] || !a->lift_out[k]) return FNSM_E_BAD_ARG;
  if (E == 0) return FNSM_OK;
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == FNSM_F64 && (!cfg || cfg->variant != 2))
    return launch_wave3d_dmma(a, E, cfg, di, st);
  // variant 2 or fp32: three back-to-back launches through the per-einsum dispatch
  const void* f1[1] = {a->v}; void* o1[1] = {a->div_out};
  int rc = opmat_dispatch(FNSM_OP_DIV, dtype, a->J, a->D, f1, o1, 1, 3, 35, 35, E, cfg, st);
  if (rc) return rc;
  const void* f2[1] = {a->u}; void* o2[1] = {a->grad_out};
  rc = opmat_dispatch(FNSM_OP_GRAD, dtype, a->J, a->D, f2, o2, 1, 3, 35, 35, E, cfg, st);
  if (rc) return rc;
  return opmat_dispatch(FNSM_OP_LIFT_FE, dtype, a->Jface, a->L, a->F, a->lift_out, 4, 4, 35, 15, E, cfg, st);
}

extern "C" int fnsm_b200_opmat_se(int32_t dtype, int32_t jac_layout, const void* const* jacs, const void* op,
                                  const void* const* fields, void* const* outs, int32_t b,
                                  int32_t n_s, int32_t n_i, int32_t n_j, int64_t E,
                                  const fnsm_cfg* cfg, void* stream) {
  using namespace fnsm;
  if (!jacs || !op || !fields || !outs || b <= 0 || E < 0) return FNSM_E_BAD_ARG;
  if (n_s < 1 || n_i < 1 || n_j < 1 || jac_layout < 0 || jac_layout > 1) return FNSM_E_BAD_ARG;
  if (cfg && (cfg->variant < 0 || cfg->variant > 3)) return FNSM_E_BAD_CONFIG;
  if (E == 0) return FNSM_OK;
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  return launch_se(dtype, jac_layout, jacs, op, fields, outs, b, n_s, n_i, n_j, E, cfg, di, static_cast<cudaStream_t>(stream));
}

extern "C" int fnsm_b200_opmat_se_supported(int32_t dtype, int32_t n_s, int32_t n_i, int32_t n_j) {
  return fnsm::se_supported(dtype, n_s, n_i, n_j) ? 1 : 0;
}
