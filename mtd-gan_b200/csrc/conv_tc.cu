// tcgen05 / TMEM implicit-GEMM convolution (TF32) — under construction in this commit: the entry
// points exist so the ABI is stable; until the kernel lands they report "unsupported" and the host
// mirror routes every layer to the exact-fp32 kernel in conv_simt.cu.
#include "common.cuh"
#include "mtdgan_b200.h"

extern "C" {
int mtd_conv_fwd_tc_supported(int, int, int, int, int, int, int, int, int, int) { return 0; }
int mtd_conv_fwd_tc(const float*, const float*, const float*, const float*, const float*, float*, float*, const float*,
                    const float*, int, int, int, int, int, int, int, int, int, int, int, int, float, void*) {
  return MTD_EINVAL;
}
}
