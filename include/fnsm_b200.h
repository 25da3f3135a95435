/*
 * fnsm_b200.h -- C ABI of the B200-native batched-einsum backend.
 *
 * This is the drop-in boundary for feinsum's execution path.  The reference
 * (kaushikcfd/feinsum 2025.3) has no native interface: a kernel is launched
 * from Python as
 *
 *     executor = t_unit.executor(cq, **arg_dict)        (src/feinsum/measure.py:163, 244)
 *     evt, outs = executor(cq, allocator=..., **arg_dict)   (measure.py:164, 251, 268)
 *
 * where t_unit comes from generate_loopy() + a transform script
 * (src/feinsum/codegen/loopy.py:112-325, src/feinsum/tuning/impls/...).  The
 * functions below are what a feinsum executor binds instead (ctypes/cffi);
 * each entry cites the reference code path it replaces.
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer owned by the caller, C-contiguous,
 *     laid out exactly as the einsum's operand shapes with the symbolic axis
 *     bound to E (reference measure.py:91-97, 226-233).  No allocation and no
 *     synchronisation happens inside; launches are asynchronous on `stream`
 *     (a cudaStream_t passed as void*), like the reference's in-order queue.
 *     Small launches of the fp64 tensor kernels carry the programmatic-stream-
 *     serialization attribute and execute griddepcontrol.wait before their first
 *     global access: every read and write keeps stream order, only launch latency
 *     and on-chip set-up overlap the previous kernel's tail.
 *   - return value: 0 on success, a positive cudaError_t, or a negative
 *     FNSM_E_* code.  fnsm_b200_strerror() decodes all three.
 *   - thread safety: concurrent calls on distinct streams are safe.  Global
 *     state: a per-device attribute cache guarded by a mutex, a relaxed atomic
 *     launch counter (fnsm_b200_launch_count), per-thread caches of encoded TMA
 *     descriptors, and the per-kernel "shared memory opted in" flags (atomics).
 *   - `cfg` may be NULL (built-in default launch configuration).  A config
 *     outside the legal space yields FNSM_E_BAD_CONFIG -- the tuner maps that
 *     to InvalidParameterError (reference tuning/__init__.py:557-559).
 */
#ifndef FNSM_B200_H
#define FNSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNSM_ABI_VERSION 1

/* dtype codes (reference: BatchedEinsum.arg_to_dtype, einsum.py:262-272) */
enum {
  FNSM_F64 = 0, FNSM_F32 = 1,
  /* generic kernel only (reference measure.py:63-77 generates ints and complex too): */
  FNSM_I32 = 2, FNSM_I64 = 3, FNSM_C64 = 4, FNSM_C128 = 5
};

/* error codes */
enum {
  FNSM_OK = 0,
  FNSM_E_BAD_ARG = -1,       /* null pointer, negative size, unknown enum  */
  FNSM_E_UNSUPPORTED = -2,   /* shape/dtype has no compiled instantiation  */
  FNSM_E_BAD_CONFIG = -3,    /* launch configuration outside legal space   */
  FNSM_E_ALIGNMENT = -4,     /* pointer not aligned for the requested path */
  FNSM_E_NO_DEVICE = -5
};

/* "operator matrix x element batch" einsum classes */
enum {
  FNSM_OP_GRAD = 0,    /* xre,rij,ej->xei   J(X,R,E) D(R,I,J) u(E,J)        -> out(X,E,I)  */
  FNSM_OP_DIV = 1,     /* xre,rij,xej->ei   J(X,R,E) D(R,I,J) u(X,E,J)      -> out(E,I)    */
  FNSM_OP_LIFT_EF = 2, /* ef,fij,fej->ei    J(E,F)   R(F,I,J) v_k(F,E,J)    -> out_k(E,I)  */
  FNSM_OP_LIFT_FE = 3  /* ifj,fe,fej->ei    L(I,F,J) J(F,E)   v_k(F,E,J)    -> out_k(E,I)  */
};

/*
 * Launch configuration: the CUDA replacement of a loopy transform script's
 * parameters (reference tuning/impls/xre_rij_ej_to_xei.py:14-38: n_e_per_wg,
 * nwork_items_per_e, i_tiles, j_tiles; xre_rij_xej_to_ei_v6.py:111-131).
 * Zero in any field = kernel default.
 */
typedef struct fnsm_cfg {
  int32_t variant;        /* 0 = auto, 1 = mma.sync tensor path (DMMA / 3xTF32), 2 = simt, 3 = tcgen05 3xTF32 (fp32) */
  int32_t tile_e;         /* elements per CTA tile                                   */
  int32_t threads;        /* threads per CTA                                         */
  int32_t stages;         /* depth of the global->shared pipeline                    */
  int32_t ctas_per_sm;    /* persistent grid = ctas_per_sm * #SM (0 = kernel default)*/
  int32_t reserved[3];    /* 0 = defaults.  Measurement / test switches of the fp64 tensor kernels, not tunables:
                           * [0] bit 0 force the non-TMA (bulk-copy) producer, bits 1-2 skip loads / stores (timing
                           * only, results invalid), bit 3 direct stores; [1] start-up phase offset between the
                           * warps of a sub-partition (cycles); [2] bits 0-3 phase-ablation build of the divergence
                           * kernel (results invalid), bits 4-5: 1 / 2 force the small-launch ("fast start")
                           * instantiations on / off instead of choosing by the number of work items per warp */
} fnsm_cfg;

/* one integer tunable and its legal range, for fnsm_b200_query_cfg_space */
typedef struct fnsm_cfg_range {
  char name[24];
  int32_t lo, hi, step, dflt;
} fnsm_cfg_range;

/* kernel ids for fnsm_b200_query_cfg_space / DB "transform_id" column */
enum {
  FNSM_K_GENERIC = 0,
  FNSM_K_GRAD = 1,
  FNSM_K_DIV = 2,
  FNSM_K_LIFT = 3,
  FNSM_K_WAVE3D = 4,
  FNSM_K_TENSOR_PRODUCT = 5,
  FNSM_K_SE = 6,
  FNSM_K_HEX_DERIV = 7
};

/* ------------------------------------------------------------------------
 * Generic batched einsum: any BatchedEinsum, one thread per output entry,
 * the reduction executed as a sequential odometer loop.  This is the CUDA
 * form of the loop nest generate_loopy() emits with the *trivial* schedule
 * (reference codegen/loopy.py:289-305) and exists so that every einsum the
 * front-end accepts executes on the GPU; the classes below are the fast paths.
 * ---------------------------------------------------------------------- */
#define FNSM_MAX_INDICES 12
#define FNSM_MAX_OPERANDS 6

typedef struct fnsm_einsum_desc {
  int32_t n_free;                         /* output rank; indices [0,n_free) are the output's, in order */
  int32_t n_sum;                          /* contracted indices [n_free, n_free+n_sum)                  */
  int32_t n_operands;
  int32_t dtype;                          /* FNSM_F64 .. FNSM_C128 (all operands and the output)        */
  int64_t extent[FNSM_MAX_INDICES];       /* symbolic extents already bound                             */
  int64_t out_stride[FNSM_MAX_INDICES];   /* element strides of the output per free index               */
  int64_t in_stride[FNSM_MAX_OPERANDS][FNSM_MAX_INDICES]; /* 0 if the operand lacks the index; summed if it repeats */
} fnsm_einsum_desc;

/* replaces: executor launch of the untransformed kernel, measure.py:163-165.
 * `inputs` = b rows x n_operands device pointers (row-major), `outputs` = b pointers. */
int fnsm_b200_generic_einsum(const fnsm_einsum_desc* desc, int32_t b,
                             const void* const* inputs, void* const* outputs,
                             void* stream);

/* ------------------------------------------------------------------------
 * Operator-matrix x element-batch einsums (DG grad / div / face-mass lift).
 * replaces: generate_loopy + tuning/impls/{xre_rij_ej_to_xei, xre_rij_xej_to_ei*,
 * ifj_fe_fej_to_ei*, batched_*}.py + executor launch (measure.py:244-273).
 *   kind      FNSM_OP_*
 *   jac       geometric factors (J), op = constant operator (D / R / L)
 *   fields    b device pointers u_k / v_k;  outs = b device pointers
 *   n_outer   ndim (grad/div: extent of x and r) or nfaces (lift: extent of f)
 *   n_i, n_j  output dofs per element, contracted dofs per element (35,35 / 35,15 for p=4 tets)
 * ---------------------------------------------------------------------- */
int fnsm_b200_opmat_batch(int32_t kind, int32_t dtype,
                          const void* jac, const void* op,
                          const void* const* fields, void* const* outs, int32_t b,
                          int32_t n_outer, int32_t n_i, int32_t n_j,
                          int64_t E, const fnsm_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------
 * Shared-operator family  se,sij,ej->ei :  out_b[e,i] = sum_{s,j} R[s,i,j] J_b[s,e] u_b[e,j]
 * replaces: generate_loopy + tuning/impls/{re_rij_ej_to_ei*, re_rji_ej_to_ei_3d_cross_product_v0}.py
 * (reference test/test_codegen.py:34-88: "div components" J{x,y,z}(3,E) R(3,35,35) u{x,y,z}(E,35),
 *  "face mass" J(4,E) R(4,15,15) v_k(E,15)).
 *   jac_layout 0: J_b(n_s, E)  "se,sij,ej->ei" (test/test_codegen.py:34-88);
 *              1: J_b(E, n_s)  "es,sij,ej->ei" (examples/dg_wave_div.py:14, test/test_feinsum.py:42)
 *   jacs     b device pointers J_b   (rows may share one array)
 *   op       R(n_s, n_i, n_j), shared by all rows
 *   fields   b device pointers u_b(E, n_j);  outs = b device pointers out_b(E, n_i)
 * Compiled shapes: fnsm_b200_opmat_se_supported(dtype, n_s, n_i, n_j) != 0 (fp64; S = 3 with 4/10/20/35 dofs,
 * S = 4 with 3/6/10/15 dofs); anything else returns FNSM_E_UNSUPPORTED and belongs to the generic kernel.
 * ---------------------------------------------------------------------- */
int fnsm_b200_opmat_se(int32_t dtype, int32_t jac_layout, const void* const* jacs, const void* op,
                       const void* const* fields, void* const* outs, int32_t b,
                       int32_t n_s, int32_t n_i, int32_t n_j, int64_t E,
                       const fnsm_cfg* cfg, void* stream);
int fnsm_b200_opmat_se_supported(int32_t dtype, int32_t n_s, int32_t n_i, int32_t n_j);

/* ------------------------------------------------------------------------
 * wave_3d_p4: div(v) + grad(u) + 4-field face-mass lift behind ONE call on one
 * stream (three kernel launches, the 2nd and 3rd with programmatic dependent
 * launch so that a kernel's tail overlaps the next prologue; the einsums share
 * only J -- 72 of 5 456 B per element -- and the operator matrices).
 * replaces: the three autotuned kernels of examples/wave_3d_p4_auto.py:16-63.
 * ---------------------------------------------------------------------- */
typedef struct fnsm_wave_args {
  const void* J;        /* (3,3,E)  */
  const void* D;        /* (3,35,35) */
  const void* v;        /* (3,E,35)  -> div_out  */
  const void* u;        /* (E,35)    -> grad_out */
  const void* L;        /* (35,4,15) */
  const void* Jface;    /* (4,E)     */
  const void* F[4];     /* (4,E,15) each */
  void* div_out;        /* (E,35)   */
  void* grad_out;       /* (3,E,35) */
  void* lift_out[4];    /* (E,35) each */
} fnsm_wave_args;

int fnsm_b200_wave3d_fused(int32_t dtype, const fnsm_wave_args* args, int64_t E,
                           const fnsm_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------
 * Tensor-product sum-factorisation on hexes: one 1-D operator applied along
 * one of the three tensor directions,
 *   mode 0: eabc,ia->eibc   mode 1: eabc,ib->eaic   mode 2: eabc,ic->eabi
 * A and out are (E, n1d, n1d, n1d); M is (n1d, n1d) row-major M[i][a].
 * (No transform exists for this in the reference)
 * ---------------------------------------------------------------------- */
int fnsm_b200_tensor_product(int32_t dtype, const void* A, const void* M, void* out,
                             int32_t n1d, int32_t mode, int64_t E,
                             const fnsm_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------
 * Fused hex derivative: the three tensor-product modes applied to the SAME field,
 *   outs[0] = eabc,ia->eibc (M[0])   outs[1] = eabc,ib->eaic (M[1])   outs[2] = eabc,ic->eabi (M[2])
 * with A(E,n1d,n1d,n1d) read once: 16 KB instead of 24 KB of traffic per fp64 p = 7 element.
 * M = 3 device pointers to (n1d, n1d) row-major operators (may be the same array), outs = 3 device pointers.
 * Compiled for fp64, n1d = 8 (FNSM_E_UNSUPPORTED otherwise); A and outs 16-byte aligned.
 * cfg->stages (2..4, default 3): elements in flight per warp; cfg->ctas_per_sm caps the persistent grid.
 * ---------------------------------------------------------------------- */
int fnsm_b200_hex_deriv(int32_t dtype, const void* A, const void* const* M, void* const* outs,
                        int32_t n1d, int64_t E, const fnsm_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------
 * Tuning-space introspection (replaces the @transform_param declarations of a
 * transform script, reference tuning/__init__.py:109-194).  Writes up to `cap`
 * ranges, returns the number of tunables of the kernel (or a negative error).
 * ---------------------------------------------------------------------- */
int fnsm_b200_query_cfg_space(int32_t kernel_id, fnsm_cfg_range* out, int32_t cap);

/* ------------------------------------------------------------------------
 * Device peaks for the roofline (replaces the static table
 * src/feinsum/data/device_info.py:4-28 with measurements on the running box).
 *   which: 0 = FP64 FMA GFLOP/s, 1 = FP32 FMA GFLOP/s, 2 = HBM copy GB/s,
 *          3 = FP64 DMMA (mma.sync m8n8k4) GFLOP/s
 *   which + 16: the SUSTAINED figure -- the micro-kernel first runs for ~0.7 s so the board sits at
 *          its power cap (FP64 tensor work pulls the B200's 1000 W limit and the SM clock drops);
 *          use it as the roofline denominator of a kernel timed inside a long step.
 * Synchronous; runs a register-resident micro-kernel for a few milliseconds (burst).
 * ---------------------------------------------------------------------- */
int fnsm_b200_measure_peak(int32_t which, double* result);

/* strided 2-D copy helper for the host-buffer path (cudaMemcpy2DAsync):
 * kind 0 = host->device, 1 = device->host.  Host memory should be pinned. */
int fnsm_b200_copy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch,
                           int64_t width_bytes, int64_t height, int32_t kind, void* stream);

/* number of kernel launches issued through this library since load (all threads) */
int64_t fnsm_b200_launch_count(void);

int fnsm_b200_abi_version(void);
const char* fnsm_b200_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif /* FNSM_B200_H */
