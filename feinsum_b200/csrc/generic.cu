// Generic batched einsum: the CUDA form of the loop nest the reference's
// generate_loopy() emits for the trivial contraction schedule
// (reference src/feinsum/codegen/loopy.py:289-305):
//     out[free...] = sum_{sum idx} prod_k operand_k[...]
// One thread per output entry (innermost output index fastest -> coalesced
// stores), reduction as an odometer over the contracted index box.  This is
// the completeness path -- any BatchedEinsum the front-end accepts runs on
// the GPU; the DG classes have dedicated kernels (opmat.cu, tensor_product.cu).
#include "common.cuh"

namespace fnsm {

// complex scalar for the generic kernel (the IR accepts complex operands, reference measure.py:63-77)
template <typename R>
struct Cx {
  R re, im;
  __device__ Cx() : re(0), im(0) {}
  __device__ Cx(int v) : re((R)v), im(0) {}
  __device__ Cx& operator*=(const Cx& o) {
    const R r = re * o.re - im * o.im;
    im = re * o.im + im * o.re;
    re = r;
    return *this;
  }
  __device__ Cx& operator+=(const Cx& o) { re += o.re; im += o.im; return *this; }
};

struct GenericRows {
  const void* in[8][FNSM_MAX_OPERANDS];
  void* out[8];
};

template <typename T>
__global__ void __launch_bounds__(256)
k_generic_einsum(const fnsm_einsum_desc d, const GenericRows rows, int nrows, long long n_out) {
  const long long tid0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const int nf = d.n_free, ns = d.n_sum, nop = d.n_operands;
  long long n_sum_iters = 1;
  for (int s = 0; s < ns; ++s) n_sum_iters *= d.extent[nf + s];

  for (long long lin = tid0; lin < n_out * nrows; lin += nthreads) {
    const int row = (int)(lin / n_out);
    long long rem = lin - (long long)row * n_out;
    // decode free indices, last output index fastest
    long long off_out = 0;
    long long base[FNSM_MAX_OPERANDS];
#pragma unroll
    for (int k = 0; k < FNSM_MAX_OPERANDS; ++k) base[k] = 0;
    for (int f = nf - 1; f >= 0; --f) {
      const long long ext = d.extent[f];
      const long long i = rem % ext;
      rem /= ext;
      off_out += i * d.out_stride[f];
      for (int k = 0; k < nop; ++k) base[k] += i * d.in_stride[k][f];
    }
    const T* __restrict__ p[FNSM_MAX_OPERANDS];
    for (int k = 0; k < nop; ++k) p[k] = static_cast<const T*>(rows.in[row][k]) + base[k];

    T acc = 0;
    int idx[FNSM_MAX_INDICES];
    for (int s = 0; s < ns; ++s) idx[s] = 0;
    long long off[FNSM_MAX_OPERANDS];
    for (int k = 0; k < nop; ++k) off[k] = 0;
    for (long long it = 0; it < n_sum_iters; ++it) {
      T prod = 1;
      for (int k = 0; k < nop; ++k) prod *= p[k][off[k]];
      acc += prod;
      // odometer increment, last contracted index fastest
      for (int s = ns - 1; s >= 0; --s) {
        const long long ext = d.extent[nf + s];
        if (++idx[s] < ext) {
          for (int k = 0; k < nop; ++k) off[k] += d.in_stride[k][nf + s];
          break;
        }
        idx[s] = 0;
        for (int k = 0; k < nop; ++k) off[k] -= (ext - 1) * d.in_stride[k][nf + s];
      }
    }
    static_cast<T*>(rows.out[row])[off_out] = acc;
  }
}

}  // namespace fnsm

extern "C" int fnsm_b200_generic_einsum(const fnsm_einsum_desc* desc, int32_t b,
                                        const void* const* inputs, void* const* outputs,
                                        void* stream) {
  using namespace fnsm;
  if (!desc || !inputs || !outputs || b <= 0) return FNSM_E_BAD_ARG;
  const int nf = desc->n_free, ns = desc->n_sum, nop = desc->n_operands;
  if (nf < 0 || ns < 0 || nf + ns > FNSM_MAX_INDICES || nop < 1 || nop > FNSM_MAX_OPERANDS)
    return FNSM_E_UNSUPPORTED;
  if (desc->dtype < FNSM_F64 || desc->dtype > FNSM_C128) return FNSM_E_UNSUPPORTED;
  long long n_out = 1;
  for (int f = 0; f < nf; ++f) {
    if (desc->extent[f] < 0) return FNSM_E_BAD_ARG;
    n_out *= desc->extent[f];
  }
  for (int s = 0; s < ns; ++s)
    if (desc->extent[nf + s] < 0) return FNSM_E_BAD_ARG;
  if (n_out == 0) return FNSM_OK;
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // rows are processed in groups of 8 (pointer table travels as a kernel parameter)
  for (int r0 = 0; r0 < b; r0 += 8) {
    const int nr = (b - r0 < 8) ? (b - r0) : 8;
    GenericRows rows{};
    for (int r = 0; r < nr; ++r) {
      for (int k = 0; k < nop; ++k) {
        rows.in[r][k] = inputs[(size_t)(r0 + r) * nop + k];
        if (!rows.in[r][k]) return FNSM_E_BAD_ARG;
      }
      rows.out[r] = outputs[r0 + r];
      if (!rows.out[r]) return FNSM_E_BAD_ARG;
    }
    const long long work = n_out * nr;
    long long blocks = (work + 255) / 256;
    const long long cap = (long long)di.sms * 16;
    if (blocks > cap) blocks = cap;
    switch (desc->dtype) {
      case FNSM_F64: k_generic_einsum<double><<<(unsigned)blocks, 256, 0, st>>>(*desc, rows, nr, n_out); break;
      case FNSM_F32: k_generic_einsum<float><<<(unsigned)blocks, 256, 0, st>>>(*desc, rows, nr, n_out); break;
      case FNSM_I32: k_generic_einsum<int><<<(unsigned)blocks, 256, 0, st>>>(*desc, rows, nr, n_out); break;
      case FNSM_I64: k_generic_einsum<long long><<<(unsigned)blocks, 256, 0, st>>>(*desc, rows, nr, n_out); break;
      case FNSM_C64: k_generic_einsum<Cx<float>><<<(unsigned)blocks, 256, 0, st>>>(*desc, rows, nr, n_out); break;
      default: k_generic_einsum<Cx<double>><<<(unsigned)blocks, 256, 0, st>>>(*desc, rows, nr, n_out); break;
    }
    if (int rc = post_launch()) return rc;
  }
  return FNSM_OK;
}
