// NHWC data-movement kernels of the discriminator decoders and the module boundaries:
//   bilinear x2 upsample (nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
//   arch/Ours/networks.py:230-260), PixelShuffle(2) (networks.py:171),
//   clip(0,1) (networks.py:1969-1970), the dropout keep-mask multiply (networks.py:417),
//   and NCHW <-> NHWC transposes used where a drop-in module is called stand-alone.
// All are pure HBM-bound streaming kernels: float4 / coalesced along the channel dimension.
#include <algorithm>
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {

inline int grid_for(size_t n, int per_block = 256, int waves = 16) {
  size_t b = (n + per_block - 1) / per_block;
  return (int)std::max<size_t>(1, std::min<size_t>(b, (size_t)mtd_sm_count() * waves));
}

// out[b, Y, X, c]: source coordinate src = max(0, (Y+0.5)/2 - 0.5), i0 = floor, i1 = min(i0+1, H-1)
__global__ void upsample2x_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C) {
  mtd_pdl_prologue();
  const int Ho = 2 * H, Wo = 2 * W;
  size_t total = (size_t)B * Ho * Wo * C;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c = (int)(i % C);
    size_t p = i / C;
    int X = (int)(p % Wo);
    p /= Wo;
    int Y = (int)(p % Ho);
    int b = (int)(p / Ho);
    float sy = fmaxf(0.f, (Y + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.f, (X + 0.5f) * 0.5f - 0.5f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    float ly = sy - y0, lx = sx - x0;
    const float* base = in + (size_t)b * H * W * C + c;
    float v00 = __ldg(base + ((size_t)y0 * W + x0) * C), v01 = __ldg(base + ((size_t)y0 * W + x1) * C);
    float v10 = __ldg(base + ((size_t)y1 * W + x0) * C), v11 = __ldg(base + ((size_t)y1 * W + x1) * C);
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

// adjoint: din[i] gathers from Y in {2i, 2i+1 (0.75)}, {max(2i-1,0), min(2i+2, 2H-1) (0.25)}
__global__ void upsample2x_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int B, int H, int W, int C) {
  mtd_pdl_prologue();
  const int Ho = 2 * H, Wo = 2 * W;
  size_t total = (size_t)B * H * W * C;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c = (int)(i % C);
    size_t p = i / C;
    int x = (int)(p % W);
    p /= W;
    int y = (int)(p % H);
    int b = (int)(p / H);
    int ys[4] = {2 * y, 2 * y + 1, max(2 * y - 1, 0), min(2 * y + 2, Ho - 1)};
    int xs[4] = {2 * x, 2 * x + 1, max(2 * x - 1, 0), min(2 * x + 2, Wo - 1)};
    const float wt[4] = {0.75f, 0.75f, 0.25f, 0.25f};
    const float* base = dout + (size_t)b * Ho * Wo * C + c;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float row = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) row = fmaf(wt[d], __ldg(base + ((size_t)ys[a] * Wo + xs[d]) * C), row);
      acc = fmaf(wt[a], row, acc);
    }
    din[i] = acc;
  }
}

// out[b, 2h+i, 2w+j, c] = in[b, h, w, 4c + 2i + j]
__global__ void pixel_shuffle2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C,
                                      int inverse) {
  mtd_pdl_prologue();
  // C = output channels; `in` has 4C channels.  inverse: scatter direction swapped (backward).
  const int Ho = 2 * H, Wo = 2 * W;
  size_t total = (size_t)B * Ho * Wo * C;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int c = (int)(i % C);
    size_t p = i / C;
    int X = (int)(p % Wo);
    p /= Wo;
    int Y = (int)(p % Ho);
    int b = (int)(p / Ho);
    size_t lo = (((size_t)b * H + (Y >> 1)) * W + (X >> 1)) * (4 * (size_t)C) + 4 * c + 2 * (Y & 1) + (X & 1);
    if (inverse) out[lo] = __ldg(in + i);      // in = d(out of shuffle) (B,2H,2W,C); out = d(in) (B,H,W,4C)
    else out[i] = __ldg(in + lo);
  }
}

// ---- float4 variants (C % 4 == 0, 16-byte aligned, < 2^31 pixels): 32-bit index arithmetic, one 16-byte access per
//      tap instead of four 4-byte ones -- the scalar kernels above spend their time in 64-bit divisions ------------
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 b) {
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}

__global__ void __launch_bounds__(256) upsample2x_fwd_v4_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int H,
                                                                int W, int C4) {
  mtd_pdl_prologue();
  const unsigned Ho = 2 * H, Wo = 2 * W;
  const unsigned total = (unsigned)B * Ho * Wo * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % C4;
    unsigned p = i / C4;
    const unsigned X = p % Wo; p /= Wo;
    const unsigned Y = p % Ho;
    const unsigned b = p / Ho;
    const float sy = fmaxf(0.f, (Y + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.f, (X + 0.5f) * 0.5f - 0.5f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float4* base = in + (size_t)b * H * W * C4 + c;
    const float4 v00 = __ldg(base + ((size_t)y0 * W + x0) * C4), v01 = __ldg(base + ((size_t)y0 * W + x1) * C4);
    const float4 v10 = __ldg(base + ((size_t)y1 * W + x0) * C4), v11 = __ldg(base + ((size_t)y1 * W + x1) * C4);
    // same expression tree as the scalar kernel: (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)
    float4 o;
    o.x = (1.f - ly) * ((1.f - lx) * v00.x + lx * v01.x) + ly * ((1.f - lx) * v10.x + lx * v11.x);
    o.y = (1.f - ly) * ((1.f - lx) * v00.y + lx * v01.y) + ly * ((1.f - lx) * v10.y + lx * v11.y);
    o.z = (1.f - ly) * ((1.f - lx) * v00.z + lx * v01.z) + ly * ((1.f - lx) * v10.z + lx * v11.z);
    o.w = (1.f - ly) * ((1.f - lx) * v00.w + lx * v01.w) + ly * ((1.f - lx) * v10.w + lx * v11.w);
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) upsample2x_bwd_v4_kernel(const float4* __restrict__ dout, float4* __restrict__ din, int B, int H,
                                                                int W, int C4) {
  mtd_pdl_prologue();
  const int Ho = 2 * H, Wo = 2 * W;
  const unsigned total = (unsigned)B * H * W * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % C4;
    unsigned p = i / C4;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const unsigned b = p / H;
    const int ys[4] = {2 * y, 2 * y + 1, max(2 * y - 1, 0), min(2 * y + 2, Ho - 1)};
    const int xs[4] = {2 * x, 2 * x + 1, max(2 * x - 1, 0), min(2 * x + 2, Wo - 1)};
    const float wt[4] = {0.75f, 0.75f, 0.25f, 0.25f};
    const float4* base = dout + (size_t)b * Ho * Wo * C4 + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int d = 0; d < 4; ++d) row = f4_fma(wt[d], __ldg(base + ((size_t)ys[a] * Wo + xs[d]) * C4), row);
      acc = f4_fma(wt[a], row, acc);
    }
    din[i] = acc;
  }
}

// One thread per (input pixel, 4 output channels): the 16 input channels 4c .. 4c+15 are one 64-byte run; a 4 x 4
// register transpose turns them into the four output pixels' float4s (and back for the inverse direction).
__global__ void __launch_bounds__(256) pixel_shuffle2_v4_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int H,
                                                                int W, int C4, int inverse) {
  mtd_pdl_prologue();
  const unsigned Wo = 2 * W;
  const unsigned total = (unsigned)B * H * W * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % C4;
    unsigned p = i / C4;
    const unsigned w = p % W; p /= W;
    const unsigned h = p % H;
    const unsigned b = p / H;
    const size_t packed = (((size_t)b * H + h) * W + w) * (4 * (size_t)C4) + 4 * c;          // float4 index, 4C-channel side
    const size_t o00 = (((size_t)b * 2 * H + 2 * h) * Wo + 2 * w) * C4 + c;                   // float4 index, C-channel side
    const size_t opos[4] = {o00, o00 + C4, o00 + (size_t)Wo * C4, o00 + (size_t)Wo * C4 + C4};   // (i, j) = (0,0) (0,1) (1,0) (1,1)
    if (!inverse) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(in + packed + k);     // v[k] = channels 4(c4+k) + {0,1,2,3} = (i,j) of out channel c4+k
      out[opos[0]] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
      out[opos[1]] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
      out[opos[2]] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
      out[opos[3]] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
    } else {
      float4 g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = __ldg(in + opos[k]);         // g[k] = d(out) at (i,j) = k, channels c4 .. c4+3
      out[packed + 0] = make_float4(g[0].x, g[1].x, g[2].x, g[3].x);
      out[packed + 1] = make_float4(g[0].y, g[1].y, g[2].y, g[3].y);
      out[packed + 2] = make_float4(g[0].z, g[1].z, g[2].z, g[3].z);
      out[packed + 3] = make_float4(g[0].w, g[1].w, g[2].w, g[3].w);
    }
  }
}

// (B,C,H,W) <-> (B,H,W,C) through a 32x32 shared tile; hw = H*W
__global__ void transpose_chw_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  mtd_pdl_prologue();
  // per batch item: in is (rows, cols) row-major, out is (cols, rows)
  __shared__ float tile[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = __ldg(in + boff + (size_t)r * cols + c);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

__global__ void clip01_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = x[i];
    y[i] = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);      // NaN propagates like torch.clip
  }
}
__global__ void clip01_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, size_t n) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = x[i];
    dx[i] = (v >= 0.f && v <= 1.f) ? dy[i] : 0.f;    // closed interval
  }
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] * b[i];
}
// out = a + b (+ c): gradient fan-in of skip tensors
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            float* __restrict__ out, size_t n) {
  mtd_pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}

}  // namespace

extern "C" {

int mtd_upsample2x_fwd(const float* in, float* out, int B, int H, int W, int C, void* stream) {
  MTD_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0);
  size_t n = (size_t)B * 4 * H * W * C;
  if (C % 4 == 0 && n < (1ull << 31) && mtd_aligned16(in) && mtd_aligned16(out)) {
    mtd_launch(upsample2x_fwd_v4_kernel, grid_for(n / 4), 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(in),
               reinterpret_cast<float4*>(out), B, H, W, C / 4);
    MTD_CHECK_LAUNCH();
    return MTD_OK;
  }
  mtd_launch(upsample2x_fwd_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, in, out, B, H, W, C);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_upsample2x_bwd(const float* dout, float* din, int B, int H, int W, int C, void* stream) {
  MTD_REQUIRE(dout && din && B > 0 && H > 0 && W > 0 && C > 0);
  size_t n = (size_t)B * H * W * C;
  if (C % 4 == 0 && n * 4 < (1ull << 31) && mtd_aligned16(dout) && mtd_aligned16(din)) {
    mtd_launch(upsample2x_bwd_v4_kernel, grid_for(n / 4), 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(dout),
               reinterpret_cast<float4*>(din), B, H, W, C / 4);
    MTD_CHECK_LAUNCH();
    return MTD_OK;
  }
  mtd_launch(upsample2x_bwd_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, dout, din, B, H, W, C);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
// forward: in (B,H,W,4C) -> out (B,2H,2W,C).  backward: in = dout (B,2H,2W,C) -> out = din (B,H,W,4C)
int mtd_pixel_shuffle2(const float* in, float* out, int B, int H, int W, int C, int backward, void* stream) {
  MTD_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0);
  size_t n = (size_t)B * 4 * H * W * C;
  if (C % 4 == 0 && n < (1ull << 31) && mtd_aligned16(in) && mtd_aligned16(out)) {
    mtd_launch(pixel_shuffle2_v4_kernel, grid_for(n / 16), 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(in),
               reinterpret_cast<float4*>(out), B, H, W, C / 4, backward);
    MTD_CHECK_LAUNCH();
    return MTD_OK;
  }
  mtd_launch(pixel_shuffle2_kernel, grid_for(n), 256, 0, (cudaStream_t)stream, in, out, B, H, W, C, backward);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
// to_nhwc != 0: in (B,C,H,W) -> out (B,H,W,C); else the inverse.
int mtd_layout_transpose(const float* in, float* out, int B, int C, int HW, int to_nhwc, void* stream) {
  MTD_REQUIRE(in && out && B > 0 && C > 0 && HW > 0 && B <= 65535);
  int rows = to_nhwc ? C : HW, cols = to_nhwc ? HW : C;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, B), block(32, 8);
  MTD_REQUIRE(grid.y <= 65535);
  mtd_launch(transpose_chw_kernel, grid, block, 0, (cudaStream_t)stream, in, out, rows, cols);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_clip01_fwd(const float* x, float* y, long long n, void* stream) {
  MTD_REQUIRE(x && y && n > 0);
  mtd_launch(clip01_fwd_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, x, y, (size_t)n);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_clip01_bwd(const float* x, const float* dy, float* dx, long long n, void* stream) {
  MTD_REQUIRE(x && dy && dx && n > 0);
  mtd_launch(clip01_bwd_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, x, dy, dx, (size_t)n);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_mul(const float* a, const float* b, float* out, long long n, void* stream) {
  MTD_REQUIRE(a && b && out && n > 0);
  mtd_launch(mul_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, a, b, out, (size_t)n);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
int mtd_add3(const float* a, const float* b, const float* c, float* out, long long n, void* stream) {
  MTD_REQUIRE(a && b && out && n > 0);
  mtd_launch(add3_kernel, grid_for((size_t)n), 256, 0, (cudaStream_t)stream, a, b, c, out, (size_t)n);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

}  // extern "C"
