// Shared device/host helpers for libmtdgan_sm100a.so.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MTD_OK 0
#define MTD_EINVAL (-1)     // bad argument / unsupported shape
#define MTD_EALIGN (-2)     // pointer not 16-byte aligned where required

// Return convention of every extern "C" entry point: 0 ok, <0 argument error,
// >0 a cudaError_t.  Nothing throws across the ABI.
// every kernel launch in the library is followed by MTD_CHECK_LAUNCH(): it also counts launches
// (mtd_kernel_launch_count) — bench.py reports that number as `gpu_launches`.
extern long long g_mtd_kernel_launches;
#define MTD_CHECK_LAUNCH()                               \
  do {                                                   \
    ++g_mtd_kernel_launches;                             \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

#define MTD_CUDA(x)                                      \
  do {                                                   \
    cudaError_t e__ = (x);                               \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

#define MTD_REQUIRE(cond)                                \
  do {                                                   \
    if (!(cond)) return MTD_EINVAL;                      \
  } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and starts with
// mtd_pdl_prologue(): `griddepcontrol.wait` blocks until the preceding grid in the stream has completed and its
// writes are visible (nothing global may be touched before it), `griddepcontrol.launch_dependents` then lets the
// NEXT grid's CTAs be scheduled (they park on their own wait), so launch latency, CTA rasterisation and each
// kernel's private setup (barrier init, TMEM allocation, descriptor prefetch) overlap the tail of the previous
// kernel instead of following it.  The trigger comes after the wait, so at most one grid runs ahead.
// A step is ~4000 launches of ~17 us: the ~1.5-2 us saved per boundary is ~10 % of the step.
extern int g_mtd_pdl;          // 1 (default) = launch with the attribute; mtd_set_pdl(0) for A/B measurements
__device__ __forceinline__ void mtd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mtd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mtd_pdl_prologue() {
  mtd_pdl_wait();
  mtd_pdl_trigger();
}

template <typename... KArgs, typename... Args>
static inline void mtd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_mtd_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface in MTD_CHECK_LAUNCH()
}

enum MtdAct { MTD_ACT_NONE = 0, MTD_ACT_RELU = 1, MTD_ACT_LEAKY = 2 };

__device__ __forceinline__ float mtd_act(float v, int act, float slope) {
  if (act == MTD_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == MTD_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}
// derivative expressed on the ACTIVATED output y (sign(y) == sign(z) for relu/leaky; at 0 both
// torch backward formulas give 0 / slope respectively — SURVEY A10).
__device__ __forceinline__ float mtd_act_grad(float y, int act, float slope) {
  if (act == MTD_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == MTD_ACT_LEAKY) return y > 0.f ? 1.f : slope;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast to all when `bcast`).  `sh` needs 32 slots.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh, bool bcast = false) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    T t = lane < nw ? sh[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) sh[0] = t;
  }
  if (bcast) {
    __syncthreads();
    v = sh[0];
  } else {
    v = (threadIdx.x == 0) ? sh[0] : T(0);
  }
  return v;
}

static inline int mtd_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline bool mtd_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
