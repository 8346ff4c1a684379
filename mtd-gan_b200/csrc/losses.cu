// Loss reductions of the MTD-GAN method wrapper, forward (fp64 block-accumulated sums) and backward
// (elementwise):  LSGAN / NDS masked squared error (losses.py:10-15), L1 restoration and RC-consistency
// MSE (arch/Ours/networks.py:1964-1977), Charbonnier (losses.py:108-111) and the Laplacian-pyramid
// EdgeLoss (losses.py:122-138).  Sums are accumulated into double slots `acc[k]` (zeroed by the
// caller) and turned into the returned float vector by mtd_loss_finalize — no host synchronisation.
//
// NDS mask: (x - y) != 0 evaluated in IEEE fp32 WITHOUT flush-to-zero (denormal -> True, -0.0 -> False,
// NaN -> True).  This translation unit must not be built with -use_fast_math / -ftz=true.
#include <algorithm>
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {

inline int grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  return (int)std::max<size_t>(1, std::min<size_t>(b, (size_t)mtd_sm_count() * 4));
}

__device__ __forceinline__ bool nds_keep(float x, float y) {
  float d = x - y;
  return (__float_as_uint(d) & 0x7fffffffu) != 0u;      // |d| != 0 on the bit pattern (NaN -> true)
}

__device__ __forceinline__ void block_accumulate(double v, double* dst) {
  __shared__ double sh[32];
  v = block_sum(v, sh);
  if (threadIdx.x == 0) atomicAdd(dst, v);
}

__global__ void nds_mask_kernel(const float* __restrict__ x, const float* __restrict__ y, unsigned char* __restrict__ m, size_t n) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) m[i] = nds_keep(x[i], y[i]) ? 1 : 0;
}

// sum mask * (in - t)^2
__global__ void sum_sqerr_kernel(const float* __restrict__ in, float t, const float* __restrict__ x, const float* __restrict__ y,
                                 size_t n, double* acc) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (; i < n; i += stride) {
    float d = in[i] - t;
    float v = d * d;
    if (x && !nds_keep(x[i], y[i])) v = 0.f;
    s += (double)v;
  }
  block_accumulate(s, acc);
}
// din = coef * mask * 2 (in - t), coef = (g[0] + g[k]) * scale
__global__ void sqerr_bwd_kernel(const float* __restrict__ in, float t, const float* __restrict__ x, const float* __restrict__ y,
                                 size_t n, const float* __restrict__ g, int k, float scale, float* __restrict__ din) {
  mtd_pdl_prologue();
  const float coef = (g[0] + g[k]) * scale * 2.f;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = coef * (in[i] - t);
    if (x && !nds_keep(x[i], y[i])) v = 0.f;
    din[i] = v;
  }
}

// mode 0: |a-b|   mode 1: (a-b)^2   mode 2: sqrt((a-b)^2 + eps^2)
__global__ void sum_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, int mode, float eps2,
                                double* acc) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (; i < n; i += stride) {
    float d = a[i] - b[i];
    float v = mode == 0 ? fabsf(d) : (mode == 1 ? d * d : sqrtf(d * d + eps2));
    s += (double)v;
  }
  block_accumulate(s, acc);
}
// da = coef * f'(a-b); db = -da (optional)
__global__ void diff_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, int mode, float eps2,
                                const float* __restrict__ g, int k, float scale, float* __restrict__ da, float* __restrict__ db) {
  mtd_pdl_prologue();
  const float coef = (g[0] + g[k]) * scale;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float d = a[i] - b[i];
    float v;
    if (mode == 0) v = d > 0.f ? coef : (d < 0.f ? -coef : 0.f);
    else if (mode == 1) v = 2.f * coef * d;
    else v = coef * d / sqrtf(d * d + eps2);
    if (da) da[i] = v;
    if (db) db[i] = -v;
  }
}

// ---- EdgeLoss: one CTA per image, everything in shared memory ---------------------------------------
__constant__ float kGauss[5] = {0.05f, 0.25f, 0.4f, 0.25f, 0.05f};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// dst = conv_gauss(src) with replicate padding; when `stuffed`, src is read as 4*src at even/even
// positions and 0 elsewhere (the zero-stuffed upsample of losses.py:129-131).
__device__ void gauss_gather(const float* src, float* dst, int H, int W, bool stuffed) {
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    int i = p / W, j = p - i * W;
    float acc = 0.f;
#pragma unroll
    for (int a = -2; a <= 2; ++a) {
      int ii = clampi(i + a, 0, H - 1);
      if (stuffed && (ii & 1)) continue;
      float row = 0.f;
#pragma unroll
      for (int b = -2; b <= 2; ++b) {
        int jj = clampi(j + b, 0, W - 1);
        if (stuffed && (jj & 1)) continue;
        row = fmaf(kGauss[b + 2], src[ii * W + jj], row);
      }
      acc = fmaf(kGauss[a + 2], row, acc);
    }
    dst[p] = stuffed ? 4.f * acc : acc;
  }
}
// adjoint of gauss_gather: dst += G^T src (dst pre-zeroed); when `stuffed`, only even/even targets
// receive (scaled by 4).
__device__ void gauss_scatter(const float* src, float* dst, int H, int W, bool stuffed) {
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    int i = p / W, j = p - i * W;
    float v = src[p];
    if (stuffed) v *= 4.f;
#pragma unroll
    for (int a = -2; a <= 2; ++a) {
      int ii = clampi(i + a, 0, H - 1);
      if (stuffed && (ii & 1)) continue;
#pragma unroll
      for (int b = -2; b <= 2; ++b) {
        int jj = clampi(j + b, 0, W - 1);
        if (stuffed && (jj & 1)) continue;
        atomicAdd(&dst[ii * W + jj], kGauss[a + 2] * kGauss[b + 2] * v);
      }
    }
  }
}

// e = lap(x - y)  (the Laplacian operator is linear, so lap(x) - lap(y) == lap(x - y))
// forward: acc += sum sqrt(e^2 + eps^2)
// backward (dfake != null): dfake = c_edge * Lap^T( e / sqrt(e^2+eps^2) ) + c_pix * d / sqrt(d^2+eps^2)
__global__ void __launch_bounds__(256) edge_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                   float eps2, double* acc, const float* __restrict__ g, int k_edge,
                                                   float scale_edge, int k_pix, float scale_pix, float* __restrict__ dfake) {
  mtd_pdl_prologue();
  extern __shared__ float sm[];
  float* d = sm;
  float* t1 = d + H * W;
  float* t2 = t1 + H * W;
  float* r = t2 + H * W;
  const size_t off = (size_t)blockIdx.x * H * W;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) d[p] = x[off + p] - y[off + p];
  __syncthreads();
  gauss_gather(d, t1, H, W, false);
  __syncthreads();
  gauss_gather(t1, t2, H, W, true);
  __syncthreads();
  if (!dfake) {
    double s = 0.0;
    for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
      float e = d[p] - t2[p];
      s += (double)sqrtf(e * e + eps2);
    }
    block_accumulate(s, acc);
    return;
  }
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    float e = d[p] - t2[p];
    r[p] = e / sqrtf(e * e + eps2);
    t1[p] = 0.f;
  }
  __syncthreads();
  // Lap^T r = r - G^T S G^T r ; forward was e = d - G(S(G d))
  gauss_scatter(r, t1, H, W, true);       // t1 = S^T-stuffed adjoint of the second conv: only even/even nonzero
  __syncthreads();
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) t2[p] = 0.f;
  __syncthreads();
  gauss_scatter(t1, t2, H, W, false);     // adjoint of the first conv
  __syncthreads();
  const float ce = (g[0] + g[k_edge]) * scale_edge;
  const float cp = (k_pix >= 0) ? (g[0] + g[k_pix]) * scale_pix : 0.f;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x) {
    float v = ce * (r[p] - t2[p]);
    if (k_pix >= 0) { float dd = d[p]; v += cp * dd / sqrtf(dd * dd + eps2); }
    dfake[off + p] = v;
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, int k, float s0, float s1, float s2, float s3,
                                     float* __restrict__ out) {
  mtd_pdl_prologue();
  const float sc[4] = {s0, s1, s2, s3};
  float total = 0.f;
  for (int i = 0; i < k; ++i) {
    float v = (float)(acc[i] * (double)sc[i]);
    out[1 + i] = v;
    total += v;        // same left-to-right fp32 association as the reference's `a + b + c + d`
  }
  out[0] = total;
}

}  // namespace

extern "C" {

int mtd_nds_mask(const float* x, const float* y, unsigned char* mask, long long n, void* stream) {
  MTD_REQUIRE(x && y && mask && n > 0);
  mtd_launch(nds_mask_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, x, y, mask, (size_t)n);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

// acc[0] += sum_i mask_i (in_i - target)^2 ; mask = NDS mask of (x, y) or all-true when x == null
int mtd_sum_sqerr(const float* in, float target, const float* x, const float* y, long long n, double* acc, void* stream) {
  MTD_REQUIRE(in && acc && n > 0 && ((x == nullptr) == (y == nullptr)));
  mtd_launch(sum_sqerr_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, in, target, x, y, (size_t)n, acc);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_sqerr_bwd(const float* in, float target, const float* x, const float* y, long long n, const float* gout, int k,
                  float scale, float* din, void* stream) {
  MTD_REQUIRE(in && gout && din && n > 0 && ((x == nullptr) == (y == nullptr)));
  mtd_launch(sqerr_bwd_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, in, target, x, y, (size_t)n, gout, k, scale, din);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
// mode 0 L1, 1 squared, 2 Charbonnier(eps)
int mtd_sum_diff(const float* a, const float* b, long long n, int mode, float eps, double* acc, void* stream) {
  MTD_REQUIRE(a && b && acc && n > 0 && mode >= 0 && mode <= 2);
  mtd_launch(sum_diff_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, a, b, (size_t)n, mode, eps * eps, acc);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_diff_bwd(const float* a, const float* b, long long n, int mode, float eps, const float* gout, int k, float scale,
                 float* da, float* db, void* stream) {
  MTD_REQUIRE(a && b && gout && (da || db) && n > 0 && mode >= 0 && mode <= 2);
  mtd_launch(diff_bwd_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, a, b, (size_t)n, mode, eps * eps, gout, k, scale, da, db);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
// EdgeLoss forward sum: acc[0] += sum sqrt(lap(x-y)^2 + eps^2) over B images of H x W
int mtd_sum_edge(const float* x, const float* y, int B, int H, int W, float eps, double* acc, void* stream) {
  MTD_REQUIRE(x && y && acc && B > 0 && H >= 3 && W >= 3);
  size_t smem = (size_t)H * W * 4 * sizeof(float);
  MTD_REQUIRE(smem <= 200 * 1024);
  if (smem > 48 * 1024) MTD_CUDA(cudaFuncSetAttribute(edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mtd_launch(edge_kernel, B, 256, smem, (cudaStream_t)stream, x, y, H, W, eps * eps, acc, nullptr, 0, 0.f, -1, 0.f, nullptr);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
// dfake = (g[0]+g[k_edge])*scale_edge * dEdge/dx  (+ (g[0]+g[k_pix])*scale_pix * dCharbonnier/dx when k_pix >= 0)
int mtd_edge_bwd(const float* x, const float* y, int B, int H, int W, float eps, const float* gout, int k_edge,
                 float scale_edge, int k_pix, float scale_pix, float* dx, void* stream) {
  MTD_REQUIRE(x && y && gout && dx && B > 0 && H >= 3 && W >= 3);
  size_t smem = (size_t)H * W * 4 * sizeof(float);
  MTD_REQUIRE(smem <= 200 * 1024);
  if (smem > 48 * 1024) MTD_CUDA(cudaFuncSetAttribute(edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mtd_launch(edge_kernel, B, 256, smem, (cudaStream_t)stream, x, y, H, W, eps * eps, nullptr, gout, k_edge, scale_edge, k_pix,
                                                      scale_pix, dx);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
// out[1+i] = acc[i]*scale[i] (i < k <= 4); out[0] = their fp32 sum
int mtd_loss_finalize(const double* acc, int k, float s0, float s1, float s2, float s3, float* out, void* stream) {
  MTD_REQUIRE(acc && out && k >= 1 && k <= 4);
  mtd_launch(loss_finalize_kernel, 1, 1, 0, (cudaStream_t)stream, acc, k, s0, s1, s2, s3, out);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

}  // extern "C"
