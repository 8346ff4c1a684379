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

__global__ void __launch_bounds__(256) adamw_kernel(const Seg* __restrict__ segs, const int2* __restrict__ chunks, float lr,
                                                    float b1, float b2, float eps, float wd) {
  mtd_pdl_prologue();
  const int2 ck = chunks[blockIdx.x];
  const Seg s = segs[ck.x];
  const long long end = min(s.numel, (long long)ck.y + kChunk);
  const float t = __ldg(s.step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step = lr / bc1, inv_sq_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  for (long long i = ck.y + threadIdx.x; i < end; i += blockDim.x) {
    float g = __ldg(s.g + i);
    float p = s.p[i] * decay;
    float m = b1 * s.m[i] + (1.f - b1) * g;
    float v = b2 * s.v[i] + (1.f - b2) * g * g;
    s.m[i] = m;
    s.v[i] = v;
    s.p[i] = p - step * m / (sqrtf(v) * inv_sq_bc2 + eps);
  }
}
}  // namespace

extern "C" int mtd_adamw_step(const void* seg_tab, int n_segs, const void* chunk_tab, int n_chunks, float lr, float beta1,
                              float beta2, float eps, float weight_decay, void* stream) {
  MTD_REQUIRE(seg_tab && chunk_tab && n_chunks > 0 && n_segs > 0);
  mtd_launch(adamw_inc_kernel, (n_segs + 127) / 128, 128, 0, (cudaStream_t)stream, reinterpret_cast<const Seg*>(seg_tab), n_segs);
  MTD_CHECK_LAUNCH();
  mtd_launch(adamw_kernel, n_chunks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const Seg*>(seg_tab),
                                                           reinterpret_cast<const int2*>(chunk_tab), lr, beta1, beta2, eps,
                                                           weight_decay);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}
