// Library-level entry points.
#include "common.cuh"
#include "mtdgan_b200.h"

long long g_mtd_kernel_launches = 0;
int g_mtd_pdl = 1;

extern "C" {

long long mtd_kernel_launch_count(void) { return g_mtd_kernel_launches; }
int mtd_set_pdl(int enabled) { int prev = g_mtd_pdl; g_mtd_pdl = enabled ? 1 : 0; return prev; }

int mtd_abi_version(void) { return 1; }

int mtd_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  return major == 10 ? 1 : 0;
}

}  // extern "C"
