// Library-wide pieces of the C ABI: error strings, device attribute cache,
// launch counter, strided host<->device copies, tuning-space introspection.
#include "common.cuh"
#include <mutex>
#include <cstring>
#include <cstdio>
#include <cstdlib>

namespace fnsm {

std::atomic<long long> g_launches{0};

static std::mutex g_dev_mu;
static DevInfo g_dev_cache[64];
static bool g_dev_valid[64] = {false};

int device_info(DevInfo* out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return FNSM_E_NO_DEVICE; }
  if (dev < 0 || dev >= 64) return FNSM_E_NO_DEVICE;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev_valid[dev]) {
    DevInfo di{};
    if (cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&di.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
      cudaGetLastError();
      return FNSM_E_NO_DEVICE;
    }
    // test hook: FNSM_B200_MAX_SMS=n makes every persistent kernel size its grid as if the device had n SMs,
    // so that small inputs walk many work items per warp (sanitizer runs, tests/test_gpu_sanitizer.py)
    if (const char* cap = std::getenv("FNSM_B200_MAX_SMS")) {
      const int v = std::atoi(cap);
      if (v > 0 && v < di.sms) di.sms = v;
    }
    g_dev_cache[dev] = di;
    g_dev_valid[dev] = true;
  }
  *out = g_dev_cache[dev];
  return FNSM_OK;
}

}  // namespace fnsm

extern "C" int fnsm_b200_abi_version(void) { return FNSM_ABI_VERSION; }

extern "C" int64_t fnsm_b200_launch_count(void) {
  return fnsm::g_launches.load(std::memory_order_relaxed);
}

extern "C" const char* fnsm_b200_strerror(int code) {
  switch (code) {
    case FNSM_OK: return "success";
    case FNSM_E_BAD_ARG: return "fnsm_b200: bad argument (null pointer, negative extent or unknown enum)";
    case FNSM_E_UNSUPPORTED: return "fnsm_b200: no compiled kernel for this shape/dtype";
    case FNSM_E_BAD_CONFIG: return "fnsm_b200: launch configuration outside the legal space";
    case FNSM_E_ALIGNMENT: return "fnsm_b200: pointer or extent not aligned for the requested kernel variant";
    case FNSM_E_NO_DEVICE: return "fnsm_b200: no usable CUDA device";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "fnsm_b200: unknown error code";
}

extern "C" int fnsm_b200_copy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch,
                                      int64_t width_bytes, int64_t height, int32_t kind, void* stream) {
  if (!dst || !src || width_bytes < 0 || height < 0 || (kind != 0 && kind != 1)) return FNSM_E_BAD_ARG;
  if (width_bytes == 0 || height == 0) return FNSM_OK;
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes,
                                    (size_t)height,
                                    kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                    static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? FNSM_OK : (int)e;
}

static void set_range(fnsm_cfg_range* r, const char* name, int lo, int hi, int step, int dflt) {
  std::memset(r, 0, sizeof(*r));
  std::snprintf(r->name, sizeof(r->name), "%s", name);
  r->lo = lo; r->hi = hi; r->step = step; r->dflt = dflt;
}

namespace fnsm { int opmat_cfg_space(int kernel_id, fnsm_cfg_range* out, int cap); }

extern "C" int fnsm_b200_query_cfg_space(int32_t kernel_id, fnsm_cfg_range* out, int32_t cap) {
  if (cap < 0 || (cap > 0 && !out)) return FNSM_E_BAD_ARG;
  fnsm_cfg_range tmp[8];
  int n = 0;
  switch (kernel_id) {
    case FNSM_K_GENERIC:
      n = 0;
      break;
    case FNSM_K_TENSOR_PRODUCT:
      set_range(&tmp[n++], "ctas_per_sm", 0, 64, 1, 0);
      break;
    case FNSM_K_SE:
      set_range(&tmp[n++], "variant", 0, 1, 1, 0);
      set_range(&tmp[n++], "ctas_per_sm", 0, 8, 1, 0);
      break;
    case FNSM_K_HEX_DERIV:
      set_range(&tmp[n++], "ctas_per_sm", 0, 8, 1, 0);
      set_range(&tmp[n++], "stages", 2, 4, 1, 3);
      break;
    case FNSM_K_GRAD: case FNSM_K_DIV: case FNSM_K_LIFT: case FNSM_K_WAVE3D:
      return fnsm::opmat_cfg_space(kernel_id, out, cap);
    default:
      return FNSM_E_BAD_ARG;
  }
  for (int i = 0; i < n && i < cap; ++i) out[i] = tmp[i];
  return n;
}
