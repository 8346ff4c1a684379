// Library-level entry points.
#include "common.cuh"
#include "mtdgan_b200.h"

extern "C" {

int mtd_abi_version(void) { return 1; }

int mtd_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  return major == 10 ? 1 : 0;
}

}  // extern "C"
