// Register-resident FFT building blocks of the Res-FFT-Conv frequency branch (fft_block.cu).
//
// A length-N transform (N = 64 ... 512) is decomposed once, N = N1 * N2 with N1, N2 <= 32 (four-step / Cooley-Tukey):
//     n = N2*n1 + n2,  k = k1 + N1*k2
//     step A   y[k1][n2] = sum_n1 x[N2*n1 + n2] W_N1^(n1 k1)          an N1-point FFT held entirely in one thread's registers
//              y[k1][n2] *= W_N^(n2 k1)                                twiddle from a shared-memory table
//     step B   X[k1 + N1*k2] = sum_n2 y[k1][n2] W_N2^(n2 k2)           an N2-point FFT in registers
// so a transform needs ONE exchange through shared memory (the transposition between the steps: written in place, read
// strided) instead of log2(N) barrier-separated radix-2 passes.  The in-register transforms are fully unrolled radix-2
// decimation-in-frequency networks (natural order in, bit-reversed order out; the index maps are compile-time constants).
#pragma once
#include <cuda_runtime.h>

namespace mtdfft {

__host__ __device__ constexpr int brev_c(int i, int n) {
  int r = 0;
  for (int b = 1, t = n >> 1; b < n; b <<= 1, t >>= 1)
    if (i & b) r |= t;
  return r;
}

// cos / sin of 2*pi*k/32, k = 0..31 (folded to immediates: every call site passes a compile-time k)
__host__ __device__ constexpr float cos32(int k) {
  k &= 31;
  if (k > 16) k = 32 - k;
  switch (k) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323044913f;
    case 2: return 0.92387953251128675613f;
    case 3: return 0.83146961230254523708f;
    case 4: return 0.70710678118654752440f;
    case 5: return 0.55557023301960222474f;
    case 6: return 0.38268343236508977173f;
    case 7: return 0.19509032201612826785f;
    case 8: return 0.0f;
    case 9: return -0.19509032201612826785f;
    case 10: return -0.38268343236508977173f;
    case 11: return -0.55557023301960222474f;
    case 12: return -0.70710678118654752440f;
    case 13: return -0.83146961230254523708f;
    case 14: return -0.92387953251128675613f;
    case 15: return -0.98078528040323044913f;
    default: return -1.0f;
  }
}
__host__ __device__ constexpr float sin32(int k) { return cos32(k - 8); }

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// In-register N-point DFT, N in {2,...,32}: v[i] <- X[brev(i)], X[k] = sum_n v[n] exp(-/+ 2 pi i n k / N)
// (INV = true: +, unnormalised).
template <int N, bool INV>
__host__ __device__ __forceinline__ void fft_reg(float2 (&v)[N]) {
  static_assert(N >= 2 && N <= 32 && (N & (N - 1)) == 0, "register FFT sizes: 2..32");
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const int pos = i % half, grp = i / half;
      const int i0 = grp * 2 * half + pos, i1 = i0 + half;
      const int tk = pos * (16 / half);                 // twiddle angle in units of 2*pi/32
      const float2 a = v[i0], b = v[i1];
      v[i0] = cadd(a, b);
      const float2 d = csub(a, b);
      if (tk == 0) {
        v[i1] = d;
      } else if (tk == 8) {                             // -i (forward) / +i (inverse)
        v[i1] = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
      } else {
        const float c = cos32(tk), s = sin32(tk);
        v[i1] = INV ? make_float2(d.x * c - d.y * s, d.y * c + d.x * s) : make_float2(d.x * c + d.y * s, d.y * c - d.x * s);
      }
    }
  }
}

}  // namespace mtdfft
