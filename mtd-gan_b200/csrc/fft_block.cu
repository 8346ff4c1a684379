// Frequency branch of the Res-FFT-Conv block, NHWC fp32, shared-memory radix-2 FFTs.
//
//   rfft2(ortho) -> cat[Re,Im] -> 1x1 conv (2C x 2C) + bias + ReLU -> complex -> irfft2(ortho)
//   (arch/Ours/networks.py:24-29), decomposed as
//     P1  fft_rows_fwd     : real FFT along W of every (b,h) row, two channels packed per complex
//                            transform; writes the half spectrum  X1[b][kw][h][c]  (complex)
//     P2  fft_cols_mix     : one CTA per (b,kw): complex FFT along H in shared memory (DIF, output
//                            left bit-reversed), per-frequency channel mix + bias + ReLU in that order,
//                            inverse FFT along H (DIT, takes bit-reversed input) -> X3[b][kw][h][c]
//     P3  fft_rows_inv     : half-spectrum inverse along W (Im of columns kw=0 and kw=W/2 dropped,
//                            SURVEY A1) fused with the block's residual adds  out = fft + add1 + add2
//   Backward: P1 on the incoming gradient, fft_cols_mix_bwd (recomputes the ReLU mask
//   from the saved X1, accumulates dW/db partials with the column weights w_k, applies M^T), then P3.
//
// No bit-reversal pass is ever executed: the channel mix is frequency-local, so it runs on the
// bit-reversed order DIF leaves behind and DIT undoes it.
#include <algorithm>
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// tw[k] = exp(-2*pi*i*k/N), k < N/2
__device__ __forceinline__ void fill_twiddles(float2* tw, int N) {
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)k / (float)N, &s, &c);
    tw[k] = make_float2(c, s);
  }
}

// Decimation-in-frequency: natural order in, bit-reversed order out.  Q interleaved sequences:
// element (n, q) lives at Z[n*Q + q].  INV uses conjugated twiddles.  Ends with __syncthreads().
template <bool INV>
__device__ __forceinline__ void fft_dif(float2* Z, const float2* tw, int N, int Q) {
  const int nb = (N >> 1) * Q;
  const int lq = __ffs(Q) - 1;                     // N, Q, half are powers of two: shifts instead of integer division
  for (int half = N >> 1; half >= 1; half >>= 1) {
    const int lh = __ffs(half) - 1;
    const int tstep = N >> (lh + 1);
    for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
      int q = idx & (Q - 1), j = idx >> lq;
      int pos = j & (half - 1), grp = j >> lh;
      int i0 = (((grp << (lh + 1)) + pos) << lq) + q, i1 = i0 + (half << lq);
      float2 a = Z[i0], b = Z[i1];
      float2 d = csub(a, b), w = tw[pos * tstep];
      Z[i0] = cadd(a, b);
      Z[i1] = INV ? cmulc(d, w) : cmul(d, w);
    }
    __syncthreads();
  }
}

// Decimation-in-time: bit-reversed order in, natural order out.
template <bool INV>
__device__ __forceinline__ void fft_dit(float2* Z, const float2* tw, int N, int Q) {
  const int nb = (N >> 1) * Q;
  const int lq = __ffs(Q) - 1;
  for (int half = 1; half < N; half <<= 1) {
    const int lh = __ffs(half) - 1;
    const int tstep = N >> (lh + 1);
    for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
      int q = idx & (Q - 1), j = idx >> lq;
      int pos = j & (half - 1), grp = j >> lh;
      int i0 = (((grp << (lh + 1)) + pos) << lq) + q, i1 = i0 + (half << lq);
      float2 w = tw[pos * tstep];
      float2 a = Z[i0], b = INV ? cmulc(Z[i1], w) : cmul(Z[i1], w);
      Z[i0] = cadd(a, b);
      Z[i1] = csub(a, b);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ int bitrev(int k, int logn) { return (int)(__brev((unsigned)k) >> (32 - logn)); }

// ------------------------------------------------------------------------------------------------
// P1: rows forward.  grid = B*H, dynamic smem = W*C*4 + (W/2)*8
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fft_rows_fwd_kernel(const float* __restrict__ x, float2* __restrict__ spec,
                                                           int H, int W, int C, int logW) {
  mtd_pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* Z = reinterpret_cast<float2*>(smem_raw);                 // [W][Q]
  const int Q = C >> 1, Wh = (W >> 1) + 1;
  float2* tw = Z + (size_t)W * Q;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  fill_twiddles(tw, W);
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)bh * W * C);
  float4* Z4 = reinterpret_cast<float4*>(Z);
  for (int i = threadIdx.x; i < W * C / 4; i += blockDim.x) Z4[i] = __ldg(src + i);
  __syncthreads();
  fft_dif<false>(Z, tw, W, Q);
  const float sc = rsqrtf((float)W) * 0.5f;
  const int lq = __ffs(Q) - 1;                      // Q is a power of two (checked by the entry point)
  for (int idx = threadIdx.x; idx < Wh * Q; idx += blockDim.x) {
    int q = idx & (Q - 1), k = idx >> lq;
    float2 zk = Z[bitrev(k, logW) * Q + q];
    float2 zm = Z[bitrev((W - k) & (W - 1), logW) * Q + q];
    zm.y = -zm.y;
    // A = (zk + zm)/2 ; B = -i (zk - zm)/2
    float4 o = make_float4((zk.x + zm.x) * sc, (zk.y + zm.y) * sc, (zk.y - zm.y) * sc, -(zk.x - zm.x) * sc);
    size_t dst = (((size_t)b * Wh + k) * H + h) * C + 2 * q;       // float2 index, even -> 16B aligned
    *reinterpret_cast<float4*>(spec + dst) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// P3: rows inverse + fused adds.  grid = B*H
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fft_rows_inv_kernel(const float2* __restrict__ spec, const float* __restrict__ add1,
                                                           const float* __restrict__ add2, float* __restrict__ out, int H,
                                                           int W, int C, int logW) {
  mtd_pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* Z = reinterpret_cast<float2*>(smem_raw);
  const int Q = C >> 1, Wh = (W >> 1) + 1;
  float2* tw = Z + (size_t)W * Q;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  fill_twiddles(tw, W);
  const int lq = __ffs(Q) - 1;
  for (int idx = threadIdx.x; idx < Wh * Q; idx += blockDim.x) {
    int q = idx & (Q - 1), k = idx >> lq;
    size_t src = (((size_t)b * Wh + k) * H + h) * C + 2 * q;
    float4 v = __ldg(reinterpret_cast<const float4*>(spec + src));   // A = (v.x, v.y), B = (v.z, v.w)
    if (k == 0 || k == (W >> 1)) {
      Z[k * Q + q] = make_float2(v.x, v.z);                          // imaginary parts dropped (A1)
    } else {
      Z[k * Q + q] = make_float2(v.x - v.w, v.y + v.z);              // A + iB
      Z[(W - k) * Q + q] = make_float2(v.x + v.w, v.z - v.y);        // conj(A) + i conj(B)
    }
  }
  __syncthreads();
  fft_dif<true>(Z, tw, W, Q);
  const float sc = rsqrtf((float)W);
  const size_t base = (size_t)bh * W * C;
  const float4* Z4 = reinterpret_cast<const float4*>(Z);
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < W * C4; i += blockDim.x) {
    int n = i >> (lq - 1), j = i & (C4 - 1);         // C4 = Q / 2
    float4 v = Z4[bitrev(n, logW) * C4 + j];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    size_t o = base + (size_t)i * 4;
    if (add1) { float4 t = __ldg(reinterpret_cast<const float4*>(add1 + o)); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    if (add2) { float4 t = __ldg(reinterpret_cast<const float4*>(add2 + o)); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    *reinterpret_cast<float4*>(out + o) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// P2: columns + channel mix (C == 32 -> 64 x 64 real mix).  grid = B*Wh
// smem: S[H][32] float2 | Mt[64][64] | bias[64] | tw[H/2]
// ------------------------------------------------------------------------------------------------
constexpr int kC = 32, kC2 = 64;

__global__ void __launch_bounds__(256) fft_cols_mix_kernel(const float2* __restrict__ spec_in, float2* __restrict__ spec_out,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           int H) {
  mtd_pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* S = reinterpret_cast<float2*>(smem_raw);
  float* Mt = reinterpret_cast<float*>(S + (size_t)H * kC);          // Mt[j][o] = w[o][j] / sqrt(H)
  float* bs = Mt + kC2 * kC2;
  float2* tw = reinterpret_cast<float2*>(bs + kC2);
  const size_t slice = (size_t)blockIdx.x * H * kC;
  const float sH = rsqrtf((float)H);
  fill_twiddles(tw, H);
  for (int i = threadIdx.x; i < kC2 * kC2; i += blockDim.x) {
    int o = i >> 6, j = i & 63;
    Mt[j * kC2 + o] = __ldg(w + i) * sH;
  }
  if (threadIdx.x < kC2) bs[threadIdx.x] = __ldg(bias + threadIdx.x);
  {
    const float4* src = reinterpret_cast<const float4*>(spec_in + slice);
    float4* S4 = reinterpret_cast<float4*>(S);
    for (int i = threadIdx.x; i < H * kC / 2; i += blockDim.x) S4[i] = __ldg(src + i);
  }
  __syncthreads();
  fft_dif<false>(S, tw, H, kC);

  const int tr = threadIdx.x >> 4, to = threadIdx.x & 15;
  float* Sw = reinterpret_cast<float*>(S);
  for (int rb = 0; rb < H; rb += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = bs[to * 4 + j];
    const int r0 = rb + tr * 4;
#pragma unroll 4
    for (int c = 0; c < kC; ++c) {
      float4 mre = *reinterpret_cast<const float4*>(Mt + c * kC2 + to * 4);
      float4 mim = *reinterpret_cast<const float4*>(Mt + (c + kC) * kC2 + to * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 y = S[(r0 + i) * kC + c];
        acc[i][0] = fmaf(y.x, mre.x, fmaf(y.y, mim.x, acc[i][0]));
        acc[i][1] = fmaf(y.x, mre.y, fmaf(y.y, mim.y, acc[i][1]));
        acc[i][2] = fmaf(y.x, mre.z, fmaf(y.y, mim.z, acc[i][2]));
        acc[i][3] = fmaf(y.x, mre.w, fmaf(y.y, mim.w, acc[i][3]));
      }
    }
    __syncwarp();      // the 16 threads sharing rows r0..r0+3 are in this warp: reads done before writes
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int o = to * 4 + j;
        Sw[((r0 + i) * kC + (o & 31)) * 2 + (o >> 5)] = fmaxf(acc[i][j], 0.f);
      }
  }
  __syncthreads();
  fft_dit<true>(S, tw, H, kC);
  {
    float4* dst = reinterpret_cast<float4*>(spec_out + slice);
    const float4* S4 = reinterpret_cast<const float4*>(S);
    for (int i = threadIdx.x; i < H * kC / 2; i += blockDim.x) {
      float4 v = S4[i];
      v.x *= sH; v.y *= sH; v.z *= sH; v.w *= sH;
      dst[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of P2.  grid = B*Wh
// smem: S[H][32] | T[H][32] | Mt[64][64] | Mn[64][64] | bias[64] | tw[H/2]
// part: per CTA 64*64 dW partial followed by 64 db partial
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fft_cols_mix_bwd_kernel(const float2* __restrict__ spec_x, const float2* __restrict__ spec_g,
                                                               float2* __restrict__ spec_out, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ part,
                                                               int H, int Wh, int W) {
  mtd_pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* S = reinterpret_cast<float2*>(smem_raw);
  float2* T = S + (size_t)H * kC;
  float* Mt = reinterpret_cast<float*>(T + (size_t)H * kC);          // Mt[j][o] = w[o][j]/sqrt(H)
  float* Mn = Mt + kC2 * kC2;                                        // Mn[o][j] = w[o][j]
  float* bs = Mn + kC2 * kC2;
  float2* tw = reinterpret_cast<float2*>(bs + kC2);
  const size_t slice = (size_t)blockIdx.x * H * kC;
  const int kw = blockIdx.x % Wh;
  const float wk = (kw == 0 || kw == (W >> 1)) ? 1.f : 2.f;
  const float sH = rsqrtf((float)H);
  fill_twiddles(tw, H);
  for (int i = threadIdx.x; i < kC2 * kC2; i += blockDim.x) {
    int o = i >> 6, j = i & 63;
    float v = __ldg(w + i);
    Mt[j * kC2 + o] = v * sH;
    Mn[i] = v;
  }
  if (threadIdx.x < kC2) bs[threadIdx.x] = __ldg(bias + threadIdx.x);
  {
    const float4* sx = reinterpret_cast<const float4*>(spec_x + slice);
    const float4* sg = reinterpret_cast<const float4*>(spec_g + slice);
    float4* S4 = reinterpret_cast<float4*>(S);
    float4* T4 = reinterpret_cast<float4*>(T);
    for (int i = threadIdx.x; i < H * kC / 2; i += blockDim.x) {
      S4[i] = __ldg(sx + i);
      T4[i] = __ldg(sg + i);
    }
  }
  __syncthreads();
  fft_dif<false>(S, tw, H, kC);
  fft_dif<false>(T, tw, H, kC);

  const int tr = threadIdx.x >> 4, to = threadIdx.x & 15;
  float* Tw = reinterpret_cast<float*>(T);
  const float* Sf = reinterpret_cast<const float*>(S);
  // 1) ReLU mask from the recomputed pre-activation; gz = mask * G2 (left unscaled) written in place
  for (int rb = 0; rb < H; rb += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = bs[to * 4 + j];
    const int r0 = rb + tr * 4;
#pragma unroll 4
    for (int c = 0; c < kC; ++c) {
      float4 mre = *reinterpret_cast<const float4*>(Mt + c * kC2 + to * 4);
      float4 mim = *reinterpret_cast<const float4*>(Mt + (c + kC) * kC2 + to * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 y = S[(r0 + i) * kC + c];
        acc[i][0] = fmaf(y.x, mre.x, fmaf(y.y, mim.x, acc[i][0]));
        acc[i][1] = fmaf(y.x, mre.y, fmaf(y.y, mim.y, acc[i][1]));
        acc[i][2] = fmaf(y.x, mre.z, fmaf(y.y, mim.z, acc[i][2]));
        acc[i][3] = fmaf(y.x, mre.w, fmaf(y.y, mim.w, acc[i][3]));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int o = to * 4 + j;
        int a = ((r0 + i) * kC + (o & 31)) * 2 + (o >> 5);
        if (!(acc[i][j] > 0.f)) Tw[a] = 0.f;
      }
  }
  __syncthreads();
  // 2) dW / db partials: dW[o][j] = wk/H * sum_r gz[r][o] * y[r][j]; db[o] = wk/sqrt(H) * sum_r gz[r][o]
  {
    const int o0 = tr * 4, j0 = to * 4;
    float acc[4][4];
    float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int r = 0; r < H; ++r) {
      float g[4], y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int o = o0 + i;
        g[i] = Tw[(r * kC + (o & 31)) * 2 + (o >> 5)];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int jj = j0 + j;
        y[j] = Sf[(r * kC + (jj & 31)) * 2 + (jj >> 5)];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        accb[i] += g[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], y[j], acc[i][j]);
      }
    }
    float* p = part + (size_t)blockIdx.x * (kC2 * kC2 + kC2);
    const float sW = wk * sH * sH;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = make_float4(acc[i][0] * sW, acc[i][1] * sW, acc[i][2] * sW, acc[i][3] * sW);
      *reinterpret_cast<float4*>(p + (o0 + i) * kC2 + j0) = v;
    }
    if (to == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) p[kC2 * kC2 + o0 + i] = accb[i] * wk * sH;
    }
  }
  __syncthreads();
  // 3) dy[r][j] = sum_o gz[r][o] * w[o][j]   (unscaled; 1/H applied at the store)  -> S
  float* Sw = reinterpret_cast<float*>(S);
  for (int rb = 0; rb < H; rb += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int r0 = rb + tr * 4;
#pragma unroll 4
    for (int o = 0; o < kC; ++o) {
      float4 mre = *reinterpret_cast<const float4*>(Mn + o * kC2 + to * 4);
      float4 mim = *reinterpret_cast<const float4*>(Mn + (o + kC) * kC2 + to * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 g = T[(r0 + i) * kC + o];     // (gz[o], gz[o+32])
        acc[i][0] = fmaf(g.x, mre.x, fmaf(g.y, mim.x, acc[i][0]));
        acc[i][1] = fmaf(g.x, mre.y, fmaf(g.y, mim.y, acc[i][1]));
        acc[i][2] = fmaf(g.x, mre.z, fmaf(g.y, mim.z, acc[i][2]));
        acc[i][3] = fmaf(g.x, mre.w, fmaf(g.y, mim.w, acc[i][3]));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int jj = to * 4 + j;
        Sw[((r0 + i) * kC + (jj & 31)) * 2 + (jj >> 5)] = acc[i][j];
      }
  }
  __syncthreads();
  fft_dit<true>(S, tw, H, kC);
  {
    float4* dst = reinterpret_cast<float4*>(spec_out + slice);
    const float4* S4 = reinterpret_cast<const float4*>(S);
    const float s2 = sH * sH;
    for (int i = threadIdx.x; i < H * kC / 2; i += blockDim.x) {
      float4 v = S4[i];
      v.x *= s2; v.y *= s2; v.z *= s2; v.w *= s2;
      dst[i] = v;
    }
  }
}

__global__ void fft_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw,
                                        float* __restrict__ db) {
  mtd_pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = kC2 * kC2 + kC2;
  if (i >= per) return;
  float s = 0.f, comp = 0.f;     // Kahan: hundreds of partials of mixed sign
  for (int p = 0; p < nparts; ++p) {
    float v = __ldg(part + (size_t)p * per + i) - comp;
    float t = s + v;
    comp = (t - s) - v;
    s = t;
  }
  if (i < kC2 * kC2) dw[i] = s;
  else db[i - kC2 * kC2] = s;
}

int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return (1 << l) == n ? l : -1;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return MTD_EINVAL;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return MTD_OK;
}

}  // namespace

extern "C" {

long long mtd_fft_spec_elems(int B, int H, int W, int C) { return (long long)B * (W / 2 + 1) * H * C * 2; }

long long mtd_fft_bwd_part_elems(int B, int W) { return (long long)B * (W / 2 + 1) * (kC2 * kC2 + kC2); }

int mtd_fft_rows_fwd(const float* x, float* spec, int B, int H, int W, int C, void* stream) {
  int lw = ilog2_exact(W);
  MTD_REQUIRE(x && spec && B > 0 && H > 0 && lw >= 3 && W <= 1024 && C >= 4 && (C & (C - 1)) == 0);    // C: power of two
  MTD_REQUIRE(mtd_aligned16(x) && mtd_aligned16(spec));
  size_t smem = (size_t)W * C * 4 + (size_t)(W / 2) * 8;
  int rc = set_smem(fft_rows_fwd_kernel, smem);
  if (rc) return rc;
  mtd_launch(fft_rows_fwd_kernel, B * H, 256, smem, (cudaStream_t)stream, x, reinterpret_cast<float2*>(spec), H, W, C, lw);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

int mtd_fft_rows_inv(const float* spec, const float* add1, const float* add2, float* out, int B, int H, int W, int C,
                     void* stream) {
  int lw = ilog2_exact(W);
  MTD_REQUIRE(spec && out && B > 0 && H > 0 && lw >= 3 && W <= 1024 && C >= 4 && (C & (C - 1)) == 0);
  MTD_REQUIRE(mtd_aligned16(spec) && mtd_aligned16(out) && mtd_aligned16(add1) && mtd_aligned16(add2));
  size_t smem = (size_t)W * C * 4 + (size_t)(W / 2) * 8;
  int rc = set_smem(fft_rows_inv_kernel, smem);
  if (rc) return rc;
  mtd_launch(fft_rows_inv_kernel, B * H, 256, smem, (cudaStream_t)stream, reinterpret_cast<const float2*>(spec), add1, add2, out,
                                                                  H, W, C, lw);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

int mtd_fft_cols_mix(const float* spec_in, float* spec_out, const float* w, const float* bias, int B, int H, int W,
                     int C, void* stream) {
  int lh = ilog2_exact(H);
  MTD_REQUIRE(spec_in && spec_out && w && bias && B > 0 && lh >= 6 && H <= 512 && W >= 2 && C == kC);
  MTD_REQUIRE(mtd_aligned16(spec_in) && mtd_aligned16(spec_out));
  size_t smem = (size_t)H * kC * 8 + (size_t)kC2 * kC2 * 4 + kC2 * 4 + (size_t)(H / 2) * 8;
  int rc = set_smem(fft_cols_mix_kernel, smem);
  if (rc) return rc;
  mtd_launch(fft_cols_mix_kernel, B * (W / 2 + 1), 256, smem, (cudaStream_t)stream, 
      reinterpret_cast<const float2*>(spec_in), reinterpret_cast<float2*>(spec_out), w, bias, H);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

int mtd_fft_cols_mix_bwd(const float* spec_x, const float* spec_g, float* spec_out, const float* w, const float* bias,
                         float* part, float* dw, float* db, int B, int H, int W, int C, void* stream) {
  int lh = ilog2_exact(H);
  MTD_REQUIRE(spec_x && spec_g && spec_out && w && bias && part && dw && db);
  MTD_REQUIRE(B > 0 && lh >= 6 && H <= 256 && W >= 2 && C == kC);
  size_t smem = (size_t)H * kC * 16 + (size_t)kC2 * kC2 * 8 + kC2 * 4 + (size_t)(H / 2) * 8;
  int rc = set_smem(fft_cols_mix_bwd_kernel, smem);
  if (rc) return rc;
  const int Wh = W / 2 + 1, nparts = B * Wh;
  cudaStream_t st = (cudaStream_t)stream;
  mtd_launch(fft_cols_mix_bwd_kernel, nparts, 256, smem, st, reinterpret_cast<const float2*>(spec_x),
                                                     reinterpret_cast<const float2*>(spec_g),
                                                     reinterpret_cast<float2*>(spec_out), w, bias, part, H, Wh, W);
  MTD_CHECK_LAUNCH();
  const int per = kC2 * kC2 + kC2;
  mtd_launch(fft_wgrad_reduce_kernel, (per + 127) / 128, 128, 0, st, part, nparts, dw, db);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

}  // extern "C"
