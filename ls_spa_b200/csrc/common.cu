// Library-wide helpers: version, status strings, cached device properties.
#include "common.cuh"

#include <stdio.h>

namespace lsspa {

const DeviceInfo &device_info() {
  static DeviceInfo info = [] {
    DeviceInfo d{0, 0, 0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return d;
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    return d;
  }();
  return info;
}

}  // namespace lsspa

extern "C" int lsspa_abi_version(void) { return LSSPA_ABI_VERSION; }

extern "C" const char *lsspa_status_string(int status) {
  static thread_local char buf[160];
  switch (status) {
    case LSSPA_OK: return "ok";
    case LSSPA_E_BADARG: return "bad argument";
    case LSSPA_E_WORKSPACE: return "workspace missing or too small";
    case LSSPA_E_NODEVICE: return "no usable CUDA device";
    case LSSPA_E_UNSUPPORTED: return "unsupported configuration";
    default: break;
  }
  if (status <= -1000) {
    cudaError_t e = (cudaError_t)(-status - 1000);
    snprintf(buf, sizeof(buf), "CUDA error %d (%s): %s", (int)e, cudaGetErrorName(e), cudaGetErrorString(e));
    return buf;
  }
  snprintf(buf, sizeof(buf), "unknown status %d", status);
  return buf;
}

extern "C" int lsspa_device_sm_count(void) { return lsspa::device_info().sm_count; }
extern "C" int lsspa_device_smem_optin(void) { return lsspa::device_info().smem_optin; }
