// Fused multi-tensor AdamW step (decoupled weight decay), one launch for a whole parameter group.
// Replaces torch.optim.AdamW.step (optimizers.py:9, train.py:122-124, engine.py:44,52): for each
// parameter with a gradient
//   p *= 1 - lr*wd ; m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ;
//   p -= lr/bias_corr1 * m / (sqrt(v)/sqrt(bias_corr2) + eps)
// Parameters whose grad is None are simply absent from the segment table.
// HBM traffic: read p,g,m,v + write p,m,v = 28 B per element.
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {
constexpr int kChunk = 16384;   // must equal the PCGrad chunk (shared host-side chunk builder)
struct Seg {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long numel;
  float* step;                    // device step counter of this parameter (incremented by adamw_inc_kernel)
  long long pad0, pad1;
};
static_assert(sizeof(Seg) == 64, "segment entry must be 8 x int64");

__global__ void adamw_inc_kernel(const Seg* __restrict__ segs, int nseg) {
  mtd_pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nseg) *segs[i].step += 1.f;
}

// lr_dev != nullptr: the learning rate is read from device memory (a captured CUDA graph then follows the
// scheduler: the host refreshes the value through a pinned copy node before each replay).  The bias corrections are
// evaluated in double like torch.optim.AdamW does on the host (1 - powf(b2, t) loses ~1e-4 relative at small t).
__global__ void __launch_bounds__(256) adamw_kernel(const Seg* __restrict__ segs, const int2* __restrict__ chunks, float lr_arg,
                                                    const float* __restrict__ lr_dev, float b1, float b2, float eps, float wd,
                                                    float grad_scale) {
  mtd_pdl_prologue();
  const int2 ck = chunks[blockIdx.x];
  const Seg s = segs[ck.x];
  const long long end = min(s.numel, (long long)ck.y + kChunk);
  const double t = (double)__ldg(s.step);
  const double lr = lr_dev ? (double)__ldg(lr_dev) : (double)lr_arg;
  const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
  const float step = (float)(lr / bc1), inv_sq_bc2 = (float)(1.0 / sqrt(bc2)), decay = (float)(1.0 - lr * (double)wd);
  const float4* g4 = reinterpret_cast<const float4*>(s.g);
  float4* p4 = reinterpret_cast<float4*>(s.p);
  float4* m4 = reinterpret_cast<float4*>(s.m);
  float4* v4 = reinterpret_cast<float4*>(s.v);
  const bool vec = ((((uintptr_t)s.p | (uintptr_t)s.g | (uintptr_t)s.m | (uintptr_t)s.v) & 15u) == 0);
  auto upd = [&](float g, float& p, float& m, float& v) {
    g *= grad_scale;
    p *= decay;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    p -= step * m / (sqrtf(v) * inv_sq_bc2 + eps);
  };
  long long i0 = ck.y;
  if (vec) {      // chunk starts are multiples of kChunk, so ck.y is float4-aligned relative to the segment base
    const long long n4 = (end - ck.y) >> 2;
    for (long long q = threadIdx.x; q < n4; q += blockDim.x) {
      const long long j = (ck.y >> 2) + q;
      const float4 g = __ldg(g4 + j);
      float4 p = p4[j], m = m4[j], v = v4[j];
      upd(g.x, p.x, m.x, v.x); upd(g.y, p.y, m.y, v.y); upd(g.z, p.z, m.z, v.z); upd(g.w, p.w, m.w, v.w);
      p4[j] = p; m4[j] = m; v4[j] = v;
    }
    i0 = ck.y + (n4 << 2);
  }
  for (long long i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float p = s.p[i], m = s.m[i], v = s.v[i];
    upd(__ldg(s.g + i), p, m, v);
    s.p[i] = p; s.m[i] = m; s.v[i] = v;
  }
}
}  // namespace

extern "C" int mtd_adamw_step(const void* seg_tab, int n_segs, const void* chunk_tab, int n_chunks, float lr, const float* lr_dev,
                              float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  MTD_REQUIRE(seg_tab && chunk_tab && n_chunks > 0 && n_segs > 0);
  mtd_launch(adamw_inc_kernel, (n_segs + 127) / 128, 128, 0, (cudaStream_t)stream, reinterpret_cast<const Seg*>(seg_tab), n_segs);
  MTD_CHECK_LAUNCH();
  mtd_launch(adamw_kernel, n_chunks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const Seg*>(seg_tab),
                                                           reinterpret_cast<const int2*>(chunk_tab), lr, lr_dev, beta1, beta2,
                                                           eps, weight_decay, grad_scale);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
