// Variant 2 ("simt") of the operator-matrix x element-batch kernels: run-time
// sized, any ndim / ndof, fp64 and fp32, plain loads.  It follows the thread
// layout of the reference's first-generation transforms (thread <-> (element,
// output dof), operator matrix resident in local memory, J-scaling hoisted out
// of the j-sum; reference tuning/impls/xre_rij_ej_to_xei.py:104-117,
// xre_rij_xej_to_ei.py, ifj_fe_fej_to_ei.py) and serves as
//   * the fallback for shapes without a tuned instantiation, and
//   * the independent on-device cross-check of the tensor-core variants.
#pragma once
#include "common.cuh"

namespace fnsm {

struct OpmatRows {
  const void* field[8];
  void* out[8];
};

// --- grad: out[x,e,i] = sum_r J[x,r,e] * (sum_j D[r,i,j] u[e,j]) ------------
template <typename T>
__global__ void __launch_bounds__(256)
k_grad_simt(const T* __restrict__ Jg, const T* __restrict__ Dg, OpmatRows rows, int nrows,
            int nd, int ni, int nj, long long E, int tile_e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* sD = reinterpret_cast<T*>(smem_raw);            // [nd][ni][nj]
  T* sU = sD + nd * ni * nj;                         // [tile_e][nj]
  T* sJ = sU + tile_e * nj;                          // [nd*nd][tile_e]
  for (int k = threadIdx.x; k < nd * ni * nj; k += blockDim.x) sD[k] = Dg[k];
  const long long ntiles = (E + tile_e - 1) / tile_e;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long e0 = tile * tile_e;
    const int ne = (int)((E - e0 < tile_e) ? (E - e0) : tile_e);
    __syncthreads();
    for (int k = threadIdx.x; k < nd * nd * tile_e; k += blockDim.x) {
      const int xr = k / tile_e, el = k - xr * tile_e;
      sJ[k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : T(0);
    }
    for (int row = 0; row < nrows; ++row) {
      const T* __restrict__ u = static_cast<const T*>(rows.field[row]);
      T* __restrict__ out = static_cast<T*>(rows.out[row]);
      if (row > 0) __syncthreads();
      for (int k = threadIdx.x; k < ne * nj; k += blockDim.x) sU[k] = u[e0 * nj + k];
      __syncthreads();
      for (int p = threadIdx.x; p < ne * ni; p += blockDim.x) {
        const int el = p / ni, i = p - el * ni;
        T acc[4] = {0, 0, 0, 0};   // nd <= 4 output components
        for (int r = 0; r < nd; ++r) {
          const T* d = sD + (r * ni + i) * nj;
          const T* uu = sU + el * nj;
          T t = 0;
          for (int j = 0; j < nj; ++j) t = fma(d[j], uu[j], t);
          for (int x = 0; x < nd; ++x) acc[x] = fma(sJ[(x * nd + r) * tile_e + el], t, acc[x]);
        }
        for (int x = 0; x < nd; ++x) out[((long long)x * E + e0) * ni + p] = acc[x];
      }
    }
  }
}

// --- div: out[e,i] = sum_{r,j} D[r,i,j] * (sum_x J[x,r,e] u[x,e,j]) ---------
template <typename T>
__global__ void __launch_bounds__(256)
k_div_simt(const T* __restrict__ Jg, const T* __restrict__ Dg, OpmatRows rows, int nrows,
           int nd, int ni, int nj, long long E, int tile_e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* sD = reinterpret_cast<T*>(smem_raw);            // [nd][ni][nj]
  T* sW = sD + nd * ni * nj;                         // [nd(r)][tile_e][nj]
  T* sJ = sW + nd * tile_e * nj;                     // [nd*nd][tile_e]
  for (int k = threadIdx.x; k < nd * ni * nj; k += blockDim.x) sD[k] = Dg[k];
  const long long ntiles = (E + tile_e - 1) / tile_e;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long e0 = tile * tile_e;
    const int ne = (int)((E - e0 < tile_e) ? (E - e0) : tile_e);
    __syncthreads();
    for (int k = threadIdx.x; k < nd * nd * tile_e; k += blockDim.x) {
      const int xr = k / tile_e, el = k - xr * tile_e;
      sJ[k] = (el < ne) ? Jg[(long long)xr * E + e0 + el] : T(0);
    }
    for (int row = 0; row < nrows; ++row) {
      const T* __restrict__ u = static_cast<const T*>(rows.field[row]);
      T* __restrict__ out = static_cast<T*>(rows.out[row]);
      __syncthreads();
      // w[r][el][j] = sum_x J[x][r][el] * u[x][e0+el][j]
      for (int k = threadIdx.x; k < ne * nj; k += blockDim.x) {
        const int el = k / nj;
        T ux[4];
        for (int x = 0; x < nd; ++x) ux[x] = u[((long long)x * E + e0) * nj + k];
        for (int r = 0; r < nd; ++r) {
          T w = 0;
          for (int x = 0; x < nd; ++x) w = fma(sJ[(x * nd + r) * tile_e + el], ux[x], w);
          sW[r * tile_e * nj + k] = w;
        }
      }
      __syncthreads();
      for (int p = threadIdx.x; p < ne * ni; p += blockDim.x) {
        const int el = p / ni, i = p - el * ni;
        T acc = 0;
        for (int r = 0; r < nd; ++r) {
          const T* d = sD + (r * ni + i) * nj;
          const T* w = sW + (r * tile_e + el) * nj;
          for (int j = 0; j < nj; ++j) acc = fma(d[j], w[j], acc);
        }
        out[e0 * ni + p] = acc;
      }
    }
  }
}

// --- lift: out_k[e,i] = sum_{f,j} Op(f,i,j) * Jf(e,f) * v_k[f,e,j] ----------
// LIFT_EF: Jf = J[e*nf+f], Op = R[(f*ni+i)*nj+j]   (ef,fij,fej->ei)
// LIFT_FE: Jf = J[f*E+e],  Op = L[(i*nf+f)*nj+j]   (ifj,fe,fej->ei)
template <typename T, bool FE>
__global__ void __launch_bounds__(256)
k_lift_simt(const T* __restrict__ Jg, const T* __restrict__ Og, OpmatRows rows, int nrows,
            int nf, int ni, int nj, long long E, int tile_e) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* sO = reinterpret_cast<T*>(smem_raw);            // [nf][ni][nj]  (always f-major here)
  T* sW = sO + nf * ni * nj;                         // [nf][tile_e][nj]
  T* sJ = sW + nf * tile_e * nj;                     // [nf][tile_e]
  for (int k = threadIdx.x; k < nf * ni * nj; k += blockDim.x) {
    const int f = k / (ni * nj), rem = k - f * ni * nj, i = rem / nj, j = rem - i * nj;
    sO[k] = FE ? Og[(i * nf + f) * nj + j] : Og[k];
  }
  const long long ntiles = (E + tile_e - 1) / tile_e;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long e0 = tile * tile_e;
    const int ne = (int)((E - e0 < tile_e) ? (E - e0) : tile_e);
    __syncthreads();
    for (int k = threadIdx.x; k < nf * tile_e; k += blockDim.x) {
      const int f = k / tile_e, el = k - f * tile_e;
      T v = 0;
      if (el < ne) v = FE ? Jg[(long long)f * E + e0 + el] : Jg[(e0 + el) * nf + f];
      sJ[k] = v;
    }
    for (int row = 0; row < nrows; ++row) {
      const T* __restrict__ v = static_cast<const T*>(rows.field[row]);
      T* __restrict__ out = static_cast<T*>(rows.out[row]);
      __syncthreads();
      for (int k = threadIdx.x; k < ne * nj; k += blockDim.x) {
        const int el = k / nj;
        for (int f = 0; f < nf; ++f)
          sW[f * tile_e * nj + k] = sJ[f * tile_e + el] * v[((long long)f * E + e0) * nj + k];
      }
      __syncthreads();
      for (int p = threadIdx.x; p < ne * ni; p += blockDim.x) {
        const int el = p / ni, i = p - el * ni;
        T acc = 0;
        for (int f = 0; f < nf; ++f) {
          const T* o = sO + (f * ni + i) * nj;
          const T* w = sW + (f * tile_e + el) * nj;
          for (int j = 0; j < nj; ++j) acc = fma(o[j], w[j], acc);
        }
        out[e0 * ni + p] = acc;
      }
    }
  }
}

}  // namespace fnsm
