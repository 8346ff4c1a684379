// Batched spectral-norm power iteration for ALL spectrally-normalised layers of the discriminator in
// four launches per D forward (instead of ~10 ATen launches per layer x 45 layers):
//   t = W^T u ; v = t / max(|t|, eps) ; s = W v ; u = s / max(|s|, eps) ; sigma = u . s
// Only (u, v, 1/sigma) are produced — W / sigma is never materialised; 1/sigma is applied in the conv
// epilogue (conv_simt.cu / conv_tc.cu `scale`).  Semantics: torch.nn.utils.spectral_norm with
// n_power_iterations=1, dim=0 (call sites arch/Ours/networks.py:181-300), SURVEY A4.
//
// layer table (device, int64[L][8]): { W ptr, u ptr, v ptr, rows, cols, u_off, v_off, p_off } where
// u_off / v_off index the packed per-call snapshot buffers (and the s workspace) and p_off the layer's slice of the
// partial-sum workspace of W^T u: [ceil(rows/32)][cols] floats.  W^T u is reduced in a FIXED order (per-CTA partials,
// then an ordered sum), so u, v and sigma are bit-reproducible from run to run and identical on every data-parallel
// rank -- no atomics, no memset.
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {

struct SnLayer {
  const float* w;
  float* u;
  float* v;
  long long rows, cols, uoff, voff, poff;
};
static_assert(sizeof(SnLayer) == 64, "layer table entry must be 8 x int64");

constexpr int kRowsPerWtu = 32;

// work item: {layer, col0, row0, 0}; 256 columns x 32 rows per CTA
__global__ void __launch_bounds__(256) sn_wtu_kernel(const SnLayer* __restrict__ tab, const int4* __restrict__ work,
                                                     float* __restrict__ t_ws) {
  mtd_pdl_prologue();
  const int4 wk = work[blockIdx.x];
  const SnLayer L = tab[wk.x];
  const long long col = wk.y + threadIdx.x;
  if (col >= L.cols) return;
  const int r1 = (int)min((long long)wk.z + kRowsPerWtu, L.rows);
  float acc = 0.f;
#pragma unroll 8
  for (int r = wk.z; r < r1; ++r) acc = fmaf(__ldg(L.w + (size_t)r * L.cols + col), __ldg(L.u + r), acc);
  t_ws[L.poff + (long long)(wk.z / kRowsPerWtu) * L.cols + col] = acc;
}

// t = W^T u: ordered sum of the row-block partials, one thread per column (work items with row0 == 0 do the work, so
// the reduction runs on as many CTAs as the layer has 256-column chunks); staged in the snapshot slot.
__global__ void __launch_bounds__(256) sn_reduce_t_kernel(const SnLayer* __restrict__ tab, const int4* __restrict__ work,
                                                          const float* __restrict__ t_ws, float* __restrict__ v_snap) {
  mtd_pdl_prologue();
  const int4 wk = work[blockIdx.x];
  if (wk.z != 0) return;
  const SnLayer L = tab[wk.x];
  const long long col = wk.y + threadIdx.x;
  if (col >= L.cols) return;
  const float* part = t_ws + L.poff + col;
  const int nrb = (int)((L.rows + kRowsPerWtu - 1) / kRowsPerWtu);
  float x = 0.f;
  for (int rb = 0; rb < nrb; ++rb) x += part[(long long)rb * L.cols];      // fixed order
  v_snap[L.voff + col] = x;
}

// one CTA per layer: v = t / max(|t|, eps)   (t staged in the snapshot slot, normalised in place)
__global__ void __launch_bounds__(256) sn_norm_v_kernel(const SnLayer* __restrict__ tab, float* __restrict__ v_snap, float eps) {
  mtd_pdl_prologue();
  __shared__ float sh[32];
  const SnLayer L = tab[blockIdx.x];
  float* t = v_snap + L.voff;
  float ss = 0.f;
  for (int k = threadIdx.x; k < L.cols; k += blockDim.x) { const float x = t[k]; ss = fmaf(x, x, ss); }
  ss = block_sum(ss, sh, true);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int k = threadIdx.x; k < L.cols; k += blockDim.x) {
    const float x = t[k] * inv;
    L.v[k] = x;
    t[k] = x;
  }
}

// work item: {layer, row0}; 8 rows per CTA, one warp per row
__global__ void __launch_bounds__(256) sn_wv_kernel(const SnLayer* __restrict__ tab, const int2* __restrict__ work,
                                                    float* __restrict__ s_ws) {
  mtd_pdl_prologue();
  const int2 wk = work[blockIdx.x];
  const SnLayer L = tab[wk.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = wk.y + warp;
  if (row >= L.rows) return;
  const float* wr = L.w + (size_t)row * L.cols;
  float acc = 0.f;
  for (long long k = lane; k < L.cols; k += 32) acc = fmaf(__ldg(wr + k), L.v[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) s_ws[L.uoff + row] = acc;
}

// one CTA per layer: (update) u = s / max(|s|, eps); sigma = u . s ; inv_sigma = 1 / sigma
__global__ void __launch_bounds__(256) sn_norm_u_kernel(const SnLayer* __restrict__ tab, const float* __restrict__ s_ws,
                                                        float* __restrict__ u_snap, float* __restrict__ inv_sigma,
                                                        int update, float eps) {
  mtd_pdl_prologue();
  __shared__ float sh[32];
  const SnLayer L = tab[blockIdx.x];
  const float* s = s_ws + L.uoff;
  float sigma;
  if (update) {
    float ss = 0.f;
    for (int r = threadIdx.x; r < L.rows; r += blockDim.x) { float x = s[r]; ss = fmaf(x, x, ss); }
    ss = block_sum(ss, sh, true);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    float dot = 0.f;
    for (int r = threadIdx.x; r < L.rows; r += blockDim.x) {
      float x = s[r] * inv;
      L.u[r] = x;
      u_snap[L.uoff + r] = x;
      dot = fmaf(x, s[r], dot);
    }
    sigma = block_sum(dot, sh, true);
  } else {
    float dot = 0.f;
    for (int r = threadIdx.x; r < L.rows; r += blockDim.x) {
      float x = L.u[r];
      u_snap[L.uoff + r] = x;
      dot = fmaf(x, s[r], dot);
    }
    sigma = block_sum(dot, sh, true);
  }
  if (threadIdx.x == 0) inv_sigma[blockIdx.x] = 1.f / sigma;
}

__global__ void sn_copy_v_kernel(const SnLayer* __restrict__ tab, float* __restrict__ v_snap) {
  mtd_pdl_prologue();
  const SnLayer L = tab[blockIdx.x];
  for (int k = threadIdx.x; k < L.cols; k += blockDim.x) v_snap[L.voff + k] = L.v[k];
}

}  // namespace

extern "C" {

int mtd_sn_rows_per_wtu_item(void) { return kRowsPerWtu; }

// update != 0: training-mode power iteration (u, v buffers updated in place, snapshots written).
// update == 0: eval mode — sigma from the stored u, v (snapshots are copies).
int mtd_sn_power_iter(const void* layer_tab, int n_layers, const void* work_wtu, int n_wtu, const void* work_wv,
                      int n_wv, float* t_ws, long long t_elems, float* s_ws, float* u_snap, float* v_snap,
                      float* inv_sigma, int update, float eps, void* stream) {
  MTD_REQUIRE(layer_tab && work_wv && s_ws && u_snap && v_snap && inv_sigma && n_layers > 0 && n_wv > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const SnLayer* tab = reinterpret_cast<const SnLayer*>(layer_tab);
  if (update) {
    MTD_REQUIRE(work_wtu && t_ws && n_wtu > 0 && t_elems > 0);
    mtd_launch(sn_wtu_kernel, n_wtu, 256, 0, st, tab, reinterpret_cast<const int4*>(work_wtu), t_ws);
    MTD_CHECK_LAUNCH();
    mtd_launch(sn_reduce_t_kernel, n_wtu, 256, 0, st, tab, reinterpret_cast<const int4*>(work_wtu), t_ws, v_snap);
    MTD_CHECK_LAUNCH();
    mtd_launch(sn_norm_v_kernel, n_layers, 256, 0, st, tab, v_snap, eps);
    MTD_CHECK_LAUNCH();
  } else {
    mtd_launch(sn_copy_v_kernel, n_layers, 256, 0, st, tab, v_snap);
    MTD_CHECK_LAUNCH();
  }
  mtd_launch(sn_wv_kernel, n_wv, 256, 0, st, tab, reinterpret_cast<const int2*>(work_wv), s_ws);
  MTD_CHECK_LAUNCH();
  mtd_launch(sn_norm_u_kernel, n_layers, 256, 0, st, tab, s_ws, u_snap, inv_sigma, update, eps);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

}  // extern "C"
