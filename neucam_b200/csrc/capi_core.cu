// capi_core.cu -- ABI version and device check.
#include "host_util.h"

extern "C" int neucam_abi_version(void) { return NEUCAM_ABI_VERSION; }

extern "C" int neucam_check_device(void) {
  int dev = 0, major = 0;
  NC_CHECK_CUDA(cudaGetDevice(&dev));
  NC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 0 : NEUCAM_ERR_ARCH;
}
