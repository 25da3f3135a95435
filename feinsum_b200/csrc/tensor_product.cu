// Tensor-product sum-factorisation on hexahedra: apply a 1-D operator M (n x n)
// along one tensor direction of A(E, n, n, n):
//   mode 0: eabc,ia->eibc    mode 1: eabc,ib->eaic    mode 2: eabc,ic->eabi
// (BASELINE config 5, p = 7 -> n = 8; the reference has no transform for it,
//  SURVEY.md section 2 -- it would run generate_loopy's trivial schedule.)
//
// Arithmetic intensity is 2n flop per 2*sizeof(T) bytes (1 flop/B for fp64,
// n = 8): a pure HBM-streaming kernel.  Design:
//   * view A as (E*P, n, Q): P = n^mode outer entries, contracted axis of
//     stride Q = n^(2-mode);
//   * modes 0/1: one thread owns one VEC-wide column (e,p,:,q..q+VEC): n
//     128-bit loads at stride Q issued back to back (n*16 B in flight per
//     thread), n*n*VEC FMAs against M, n 128-bit stores.  Consecutive threads
//     own consecutive q, so every warp request covers whole 32 B sectors and,
//     for mode 0, 512 contiguous bytes;
//   * mode 2 (contracted axis contiguous): one thread owns one row of n;
//   * M (<= 8x8) is staged once per CTA in shared memory and read with
//     warp-uniform (broadcast) loads;
//   * all global traffic is touch-once: ld.global.cs / st.global.cs.
#include "common.cuh"

namespace fnsm {

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> load_pack(const T* p) {
  Pack<T, VEC> r;
  if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 8) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p)); r.v[0] = t.x; r.v[1] = t.y;
  } else if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 4) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (sizeof(T) * VEC == 8 && sizeof(T) == 4) {
    float2 t = __ldcs(reinterpret_cast<const float2*>(p)); r.v[0] = t.x; r.v[1] = t.y;
  } else {
    static_assert(VEC == 1, "unsupported pack");
    r.v[0] = __ldcs(p);
  }
  return r;
}
template <typename T, int VEC>
__device__ __forceinline__ void store_pack(T* p, const Pack<T, VEC>& r) {
  if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 8) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(r.v[0], r.v[1]));
  } else if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  } else if constexpr (sizeof(T) * VEC == 8 && sizeof(T) == 4) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(r.v[0], r.v[1]));
  } else {
    __stcs(p, r.v[0]);
  }
}

// modes 0 and 1: strided columns, VEC-wide
template <typename T, int N, int Q, int VEC>
__global__ void __launch_bounds__(256)
k_tp_col(const T* __restrict__ A, const T* __restrict__ M, T* __restrict__ out, long long n_items) {
  __shared__ T sM[N * N];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) sM[i] = M[i];
  __syncthreads();
  constexpr int QV = Q / VEC;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x; item < n_items; item += stride) {
    const long long ep = item / QV;
    const int qv = (int)(item - ep * QV);
    const long long base = ep * (long long)(N * Q) + qv * VEC;
    Pack<T, VEC> in[N];
#pragma unroll
    for (int a = 0; a < N; ++a) in[a] = load_pack<T, VEC>(A + base + a * Q);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      Pack<T, VEC> acc;
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc.v[v] = 0;
#pragma unroll
      for (int a = 0; a < N; ++a) {
        const T m = sM[i * N + a];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc.v[v] = fma(m, in[a].v[v], acc.v[v]);
      }
      store_pack<T, VEC>(out + base + i * Q, acc);
    }
  }
}

// mode 2: contiguous rows of N
template <typename T, int N, int VEC>
__global__ void __launch_bounds__(256)
k_tp_row(const T* __restrict__ A, const T* __restrict__ M, T* __restrict__ out, long long n_rows) {
  __shared__ T sM[N * N];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) sM[i] = M[i];
  __syncthreads();
  constexpr int NV = N / VEC;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    const long long base = row * N;
    T in[N];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      Pack<T, VEC> t = load_pack<T, VEC>(A + base + c * VEC);
#pragma unroll
      for (int v = 0; v < VEC; ++v) in[c * VEC + v] = t.v[v];
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      Pack<T, VEC> o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        T acc = 0;
#pragma unroll
        for (int a = 0; a < N; ++a) acc = fma(sM[(c * VEC + v) * N + a], in[a], acc);
        o.v[v] = acc;
      }
      store_pack<T, VEC>(out + base + c * VEC, o);
    }
  }
}

static inline unsigned grid_for(long long items, const DevInfo& di, const fnsm_cfg* cfg) {
  long long blocks = (items + 255) / 256;
  if (cfg && cfg->ctas_per_sm > 0) {
    const long long cap = (long long)cfg->ctas_per_sm * di.sms;
    if (blocks > cap) blocks = cap;
  }
  if (blocks > 0x7fffffffLL) blocks = 0x7fffffffLL;
  return (unsigned)blocks;
}

template <typename T, int N>
static int launch_tp(const T* A, const T* M, T* out, int mode, long long E, bool aligned16,
                     const fnsm_cfg* cfg, cudaStream_t st) {
  DevInfo di;
  if (int rc = device_info(&di)) return rc;
  constexpr int V16 = 16 / (int)sizeof(T);   // elements per 128-bit access
  if (mode == 0 || mode == 1) {
    constexpr int Q0 = N * N, Q1 = N;
    if (mode == 0) {
      if (aligned16 && (Q0 % V16) == 0) {
        const long long items = E * (Q0 / V16);
        k_tp_col<T, N, Q0, V16><<<grid_for(items, di, cfg), 256, 0, st>>>(A, M, out, items);
      } else {
        const long long items = E * Q0;
        k_tp_col<T, N, Q0, 1><<<grid_for(items, di, cfg), 256, 0, st>>>(A, M, out, items);
      }
    } else {
      if (aligned16 && (Q1 % V16) == 0) {
        const long long items = E * N * (Q1 / V16);
        k_tp_col<T, N, Q1, (Q1 % V16) == 0 ? V16 : 1><<<grid_for(items, di, cfg), 256, 0, st>>>(A, M, out, items);
      } else {
        const long long items = E * N * Q1;
        k_tp_col<T, N, Q1, 1><<<grid_for(items, di, cfg), 256, 0, st>>>(A, M, out, items);
      }
    }
  } else {
    const long long rows = E * N * N;
    if (aligned16 && (N % V16) == 0)
      k_tp_row<T, N, (N % V16) == 0 ? V16 : 1><<<grid_for(rows, di, cfg), 256, 0, st>>>(A, M, out, rows);
    else
      k_tp_row<T, N, 1><<<grid_for(rows, di, cfg), 256, 0, st>>>(A, M, out, rows);
  }
  return post_launch();
}

template <typename T>
static int dispatch_tp(const void* A, const void* M, void* out, int n1d, int mode, long long E,
                       const fnsm_cfg* cfg, cudaStream_t st) {
  const bool aligned16 = (((uintptr_t)A | (uintptr_t)out) & 15) == 0;
  const T* a = static_cast<const T*>(A);
  const T* m = static_cast<const T*>(M);
  T* o = static_cast<T*>(out);
  switch (n1d) {
    case 2: return launch_tp<T, 2>(a, m, o, mode, E, aligned16, cfg, st);
    case 3: return launch_tp<T, 3>(a, m, o, mode, E, aligned16, cfg, st);
    case 4: return launch_tp<T, 4>(a, m, o, mode, E, aligned16, cfg, st);
    case 5: return launch_tp<T, 5>(a, m, o, mode, E, aligned16, cfg, st);
    case 6: return launch_tp<T, 6>(a, m, o, mode, E, aligned16, cfg, st);
    case 7: return launch_tp<T, 7>(a, m, o, mode, E, aligned16, cfg, st);
    case 8: return launch_tp<T, 8>(a, m, o, mode, E, aligned16, cfg, st);
    default: return FNSM_E_UNSUPPORTED;
  }
}

}  // namespace fnsm

extern "C" int fnsm_b200_tensor_product(int32_t dtype, const void* A, const void* M, void* out,
                                        int32_t n1d, int32_t mode, int64_t E,
                                        const fnsm_cfg* cfg, void* stream) {
  using namespace fnsm;
  if (!A || !M || !out || E < 0 || mode < 0 || mode > 2) return FNSM_E_BAD_ARG;
  if (cfg && (cfg->ctas_per_sm < 0 || cfg->ctas_per_sm > 64)) return FNSM_E_BAD_CONFIG;
  if (E == 0) return FNSM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == FNSM_F64) return dispatch_tp<double>(A, M, out, n1d, mode, E, cfg, st);
  if (dtype == FNSM_F32) return dispatch_tp<float>(A, M, out, n1d, mode, E, cfg, st);
  return FNSM_E_UNSUPPORTED;
}
