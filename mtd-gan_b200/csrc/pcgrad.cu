// PCGrad in Gram space (module/weight_methods.py:449-464 and module/pcgrad.py:50-70).
//
// Every progressively projected g_i' stays in span{g_0..g_{T-1}}, so the whole projection loop is a
// T x T problem once the Gram matrix is known:
//   pass 1  pcgrad_gram     : one read of the T task gradients -> T(T+1)/2 dots, fp64 accumulation
//   solve   pcgrad_solve    : 1 thread, replays the reference's loop (visit orders come from the host's
//                             `random.shuffle`, passed as data — no host sync on `if dot < 0`)
//   pass 2  pcgrad_combine  : merged = scale_seg * sum_k coef[k] * g_k, one read of the T gradients +
//                             one write
// Algorithmic traffic (T=3): (3 + 3 + 1) * 4 B per element.  Gradients are addressed through a segment
// table (one segment per parameter tensor) so nothing is flattened or copied first.
//
// segment table (device int64[nseg][8]): { g_0, g_1, g_2, g_3 (0 = task has no grad: zeros), out, numel,
//                                         float-bits of scale, 0 }
// chunk table (device int32[nchunk][2]): { segment, element offset }, chunk = kChunk elements
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {

constexpr int kMaxTasks = 4;
constexpr int kChunk = 16384;

struct Seg {
  const float* g[kMaxTasks];
  float* out;
  long long numel;
  long long scale_bits;
  long long pad;
};
static_assert(sizeof(Seg) == 64, "segment entry must be 8 x int64");

__global__ void __launch_bounds__(256) pcgrad_gram_kernel(const Seg* __restrict__ segs, const int2* __restrict__ chunks,
                                                          int T, double* __restrict__ gram) {
  mtd_pdl_prologue();
  __shared__ double sh[32];
  const int2 ck = chunks[blockIdx.x];
  const Seg s = segs[ck.x];
  const long long end = min(s.numel, (long long)ck.y + kChunk);
  float acc[kMaxTasks * (kMaxTasks + 1) / 2];
#pragma unroll
  for (int i = 0; i < kMaxTasks * (kMaxTasks + 1) / 2; ++i) acc[i] = 0.f;
  double dacc[kMaxTasks * (kMaxTasks + 1) / 2];
#pragma unroll
  for (int i = 0; i < kMaxTasks * (kMaxTasks + 1) / 2; ++i) dacc[i] = 0.0;
  int since = 0;
  auto accumulate = [&](const float (&v)[kMaxTasks]) {
    int p = 0;
#pragma unroll
    for (int a = 0; a < kMaxTasks; ++a)
#pragma unroll
      for (int b = a; b < kMaxTasks; ++b) acc[p] = fmaf(v[a], v[b], acc[p]), ++p;
  };
  auto flush = [&]() {       // flush short fp32 runs into fp64 (task-2 norms are ~1e-5 of task-0's)
#pragma unroll
    for (int q = 0; q < kMaxTasks * (kMaxTasks + 1) / 2; ++q) dacc[q] += (double)acc[q], acc[q] = 0.f;
  };
  bool vec = true;
#pragma unroll
  for (int t = 0; t < kMaxTasks; ++t)
    if (t < T && s.g[t] && (((uintptr_t)s.g[t]) & 15u)) vec = false;
  long long i0 = ck.y;
  if (vec) {      // 16-byte loads: four elements of every task gradient per iteration (chunk starts are multiples of 4)
    const long long n4 = (end - ck.y) >> 2;
    for (long long q = threadIdx.x; q < n4; q += blockDim.x) {
      float4 v4[kMaxTasks];
#pragma unroll
      for (int t = 0; t < kMaxTasks; ++t)
        v4[t] = (t < T && s.g[t]) ? __ldg(reinterpret_cast<const float4*>(s.g[t] + ck.y) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      float v[kMaxTasks];
#pragma unroll
      for (int t = 0; t < kMaxTasks; ++t) v[t] = v4[t].x;
      accumulate(v);
#pragma unroll
      for (int t = 0; t < kMaxTasks; ++t) v[t] = v4[t].y;
      accumulate(v);
#pragma unroll
      for (int t = 0; t < kMaxTasks; ++t) v[t] = v4[t].z;
      accumulate(v);
#pragma unroll
      for (int t = 0; t < kMaxTasks; ++t) v[t] = v4[t].w;
      accumulate(v);
      if (++since == 2) { since = 0; flush(); }
    }
    i0 = ck.y + (n4 << 2);
    since = 0;
    flush();
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float v[kMaxTasks];
#pragma unroll
    for (int t = 0; t < kMaxTasks; ++t) v[t] = (t < T && s.g[t]) ? __ldg(s.g[t] + i) : 0.f;
    accumulate(v);
    if (++since == 8) { since = 0; flush(); }
  }
#pragma unroll
  for (int q = 0; q < kMaxTasks * (kMaxTasks + 1) / 2; ++q) dacc[q] += (double)acc[q];
  int p = 0;
  for (int a = 0; a < kMaxTasks; ++a)
    for (int b = a; b < kMaxTasks; ++b, ++p) {
      if (b >= T) continue;            // uniform
      double r = block_sum(dacc[p], sh);
      if (threadIdx.x == 0) atomicAdd(gram + a * kMaxTasks + b, r);
    }
}

// orders: int32[T][T], orders[i][*] = the task indices j in the order the reference visits them for
// outer task i.  coef_out[k] = sum_i C[i][k] * (mean ? 1/T : 1); cmat_out[T][T] = C (for tests).
__global__ void pcgrad_solve_kernel(double* __restrict__ gram, const int* __restrict__ orders, int T, int mean,
                                    float* __restrict__ coef_out, float* __restrict__ cmat_out, double* __restrict__ gram_out) {
  mtd_pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double G[kMaxTasks][kMaxTasks], C[kMaxTasks][kMaxTasks];
  for (int a = 0; a < T; ++a)
    for (int b = a; b < T; ++b) G[a][b] = G[b][a] = gram[a * kMaxTasks + b];
  for (int i = 0; i < T; ++i)
    for (int k = 0; k < T; ++k) C[i][k] = (i == k) ? 1.0 : 0.0;
  for (int i = 0; i < T; ++i)
    for (int s = 0; s < T; ++s) {
      int j = orders[i * T + s];
      double dot = 0.0;
      for (int k = 0; k < T; ++k) dot += C[i][k] * G[k][j];
      if (dot < 0.0) C[i][j] -= dot / G[j][j];
    }
  for (int k = 0; k < T; ++k) {
    double w = 0.0;
    for (int i = 0; i < T; ++i) w += C[i][k];
    coef_out[k] = (float)(mean ? w / T : w);
  }
  if (cmat_out)
    for (int i = 0; i < T; ++i)
      for (int k = 0; k < T; ++k) cmat_out[i * T + k] = (float)C[i][k];
  if (gram_out)
    for (int a = 0; a < T; ++a)
      for (int b = 0; b < T; ++b) gram_out[a * T + b] = G[a][b];
}

__global__ void __launch_bounds__(256) pcgrad_combine_kernel(const Seg* __restrict__ segs, const int2* __restrict__ chunks,
                                                             int T, const float* __restrict__ coef) {
  mtd_pdl_prologue();
  const int2 ck = chunks[blockIdx.x];
  const Seg s = segs[ck.x];
  const long long end = min(s.numel, (long long)ck.y + kChunk);
  const float scale = __int_as_float((int)s.scale_bits);
  float c[kMaxTasks];
#pragma unroll
  for (int t = 0; t < kMaxTasks; ++t) c[t] = (t < T) ? __ldg(coef + t) * scale : 0.f;
  bool vec = (((uintptr_t)s.out) & 15u) == 0;
#pragma unroll
  for (int t = 0; t < kMaxTasks; ++t)
    if (t < T && s.g[t] && (((uintptr_t)s.g[t]) & 15u)) vec = false;
  long long i0 = ck.y;
  if (vec) {
    const long long n4 = (end - ck.y) >> 2;
    for (long long q = threadIdx.x; q < n4; q += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < kMaxTasks; ++t)
        if (t < T && s.g[t]) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(s.g[t] + ck.y) + q);
          v.x = fmaf(c[t], g.x, v.x); v.y = fmaf(c[t], g.y, v.y); v.z = fmaf(c[t], g.z, v.z); v.w = fmaf(c[t], g.w, v.w);
        }
      reinterpret_cast<float4*>(s.out + ck.y)[q] = v;
    }
    i0 = ck.y + (n4 << 2);
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxTasks; ++t)
      if (t < T && s.g[t]) v = fmaf(c[t], __ldg(s.g[t] + i), v);
    s.out[i] = v;
  }
}

// out = scale_seg * g_0 for every segment: gathers a list of gradient tensors into one flat buffer (the operand of a
// reduce-scatter / all-reduce) in ONE launch, with the 1/world_size of the batch mean folded in.
__global__ void __launch_bounds__(256) segs_scale_copy_kernel(const Seg* __restrict__ segs, const int2* __restrict__ chunks) {
  mtd_pdl_prologue();
  const int2 ck = chunks[blockIdx.x];
  const Seg s = segs[ck.x];
  const long long end = min(s.numel, (long long)ck.y + kChunk);
  const float scale = __int_as_float((int)s.scale_bits);
  const float* g = s.g[0];
  long long i0 = ck.y;
  if ((((uintptr_t)g | (uintptr_t)s.out) & 15u) == 0) {
    const long long n4 = (end - ck.y) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g + ck.y);
    float4* o4 = reinterpret_cast<float4*>(s.out + ck.y);
    for (long long q = threadIdx.x; q < n4; q += blockDim.x) {
      float4 v = __ldg(g4 + q);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      o4[q] = v;
    }
    i0 = ck.y + (n4 << 2);
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) s.out[i] = scale * __ldg(g + i);
}

}  // namespace

extern "C" {

int mtd_pcgrad_chunk_elems(void) { return kChunk; }

int mtd_segments_scale_copy(const void* seg_tab, const void* chunk_tab, int n_chunks, void* stream) {
  MTD_REQUIRE(seg_tab && chunk_tab && n_chunks > 0);
  mtd_launch(segs_scale_copy_kernel, n_chunks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const Seg*>(seg_tab),
             reinterpret_cast<const int2*>(chunk_tab));
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

// gram_ws: 16 doubles (zeroed here); entry [a*4+b], a <= b.  Split form for multi-GPU use: the caller
// all-reduces gram_ws between the two calls (each rank holds a shard of the gradients).
int mtd_pcgrad_gram(const void* seg_tab, const void* chunk_tab, int n_chunks, int T, double* gram_ws, void* stream) {
  MTD_REQUIRE(seg_tab && chunk_tab && gram_ws && n_chunks > 0 && T >= 1 && T <= kMaxTasks);
  cudaStream_t st = (cudaStream_t)stream;
  MTD_CUDA(cudaMemsetAsync(gram_ws, 0, 16 * sizeof(double), st));
  mtd_launch(pcgrad_gram_kernel, n_chunks, 256, 0, st, reinterpret_cast<const Seg*>(seg_tab),
                                               reinterpret_cast<const int2*>(chunk_tab), T, gram_ws);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

// coef_out: T floats.  cmat_out (T*T floats) / gram_out (T*T doubles) may be null.
int mtd_pcgrad_solve_combine(const void* seg_tab, const void* chunk_tab, int n_chunks, int T, const int* orders, int mean,
                             double* gram_ws, float* coef_out, float* cmat_out, double* gram_out, void* stream) {
  MTD_REQUIRE(seg_tab && chunk_tab && orders && gram_ws && coef_out && n_chunks > 0 && T >= 1 && T <= kMaxTasks);
  cudaStream_t st = (cudaStream_t)stream;
  mtd_launch(pcgrad_solve_kernel, 1, 32, 0, st, gram_ws, orders, T, mean, coef_out, cmat_out, gram_out);
  MTD_CHECK_LAUNCH();
  mtd_launch(pcgrad_combine_kernel, n_chunks, 256, 0, st, reinterpret_cast<const Seg*>(seg_tab),
                                                  reinterpret_cast<const int2*>(chunk_tab), T, coef_out);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

int mtd_pcgrad_project(const void* seg_tab, const void* chunk_tab, int n_chunks, int T, const int* orders, int mean,
                       double* gram_ws, float* coef_out, float* cmat_out, double* gram_out, void* stream) {
  int rc = mtd_pcgrad_gram(seg_tab, chunk_tab, n_chunks, T, gram_ws, stream);
  if (rc) return rc;
  return mtd_pcgrad_solve_combine(seg_tab, chunk_tab, n_chunks, T, orders, mean, gram_ws, coef_out, cmat_out, gram_out, stream);
}

}  // extern "C"
