// Data edge of the training loop: the reference's `window_patch` transform chain
// (create_datasets/Mayo.py:117-136: ScaleIntensityRanged HU window [-160, 240] -> [0, 1] with clip, CropForegroundd on
// the full-dose slice, SpatialPadd to 64 x 64, RandSpatialCropSamplesd 8 x 64 x 64, RandRotate90d, RandFlipd) on the GPU,
// from int16 HU slices resident in HBM.  Byte / index work: one pass over the slices for the bounding boxes, one gather
// pass that windows only the 8 x 64 x 64 pixels per slice that are actually used.  Random decisions are made on the
// host (numpy RandomState, same draw order as the numpy restatement in oracle/data_edge.py) and arrive as a table.
#include "common.cuh"
#include "mtdgan_b200.h"

namespace {

// bbox[s] = {y0, y1, x0, x1} (half-open) of {hu > a_min} in slice s; {0,0,0,0} when the slice has no foreground.
// One CTA per slice: rows are scanned with coalesced 16-bit loads, per-thread extents reduced through shared atomics.
__global__ void __launch_bounds__(256) hu_bbox_kernel(const short* __restrict__ hu, int H, int W, float a_min, int* __restrict__ bbox) {
  mtd_pdl_prologue();
  __shared__ int sh[4];
  if (threadIdx.x == 0) { sh[0] = H; sh[1] = -1; sh[2] = W; sh[3] = -1; }
  __syncthreads();
  const short* img = hu + (size_t)blockIdx.x * H * W;
  int y0 = H, y1 = -1, x0 = W, x1 = -1;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    if ((float)img[i] > a_min) {
      const int y = i / W, x = i - y * W;
      y0 = min(y0, y); y1 = max(y1, y); x0 = min(x0, x); x1 = max(x1, x);
    }
  }
  if (y1 >= 0) { atomicMin(&sh[0], y0); atomicMax(&sh[1], y1); atomicMin(&sh[2], x0); atomicMax(&sh[3], x1); }
  __syncthreads();
  if (threadIdx.x < 4) {
    const bool empty = sh[1] < 0;
    int v = sh[threadIdx.x];
    if (threadIdx.x & 1) v += 1;                      // half-open upper bounds
    bbox[blockIdx.x * 4 + threadIdx.x] = empty ? 0 : v;
  }
}

struct PatchRow { int slice, oy, ox, by0, by1, bx0, bx1, aug; };      // aug: bits 0-1 = rot90 count k, bit 2 = flip both axes
static_assert(sizeof(PatchRow) == 32, "patch table row must be 8 x int32");

// ScaleIntensityRange(a_min, a_max, 0, 1, clip=True) evaluated like monai==1.3.2 (requirements.txt:14) does: the int16
// image becomes a torch tensor, `(img - a_min) / (a_max - a_min)` promotes to float32 (Python-float scalars), then
// `* (b_max - b_min) + b_min` and the clip (monai/transforms/intensity/array.py ScaleIntensityRange.__call__).  IEEE fp32
// subtraction and division (no fast-math), so the result is bit-identical to the torch-CPU evaluation.
__device__ __forceinline__ float hu_window(short v, float a_min, float range) {
  float t = __fdiv_rn(__fsub_rn((float)v, a_min), range);
  t = __fadd_rn(__fmul_rn(t, 1.0f), 0.0f);
  return fminf(fmaxf(t, 0.0f), 1.0f);
}

// out[n][i][j] for both dose levels.  (i, j) is mapped back through flip and rot90 to the crop pixel (pi, pj), then to the
// slice pixel (oy + pi, ox + pj); pixels outside the foreground box are SpatialPad zeros.
__global__ void __launch_bounds__(256) window_crop_kernel(const short* __restrict__ lo, const short* __restrict__ hi, int H, int W,
                                                          const PatchRow* __restrict__ tab, int roi, float a_min, float range,
                                                          float* __restrict__ x, float* __restrict__ y) {
  mtd_pdl_prologue();
  const PatchRow r = tab[blockIdx.x];
  const int k = r.aug & 3, flip = (r.aug >> 2) & 1;
  const size_t base = (size_t)r.slice * H * W;
  for (int p = threadIdx.x; p < roi * roi; p += blockDim.x) {
    int i = p / roi, j = p - i * roi;
    if (flip) { i = roi - 1 - i; j = roi - 1 - j; }                     // np.flip over both spatial axes
    // np.rot90(a, k)[i, j]: k=1 -> a[j, n-1-i]; k=2 -> a[n-1-i, n-1-j]; k=3 -> a[n-1-j, i]
    int pi = i, pj = j;
    if (k == 1) { pi = j; pj = roi - 1 - i; }
    else if (k == 2) { pi = roi - 1 - i; pj = roi - 1 - j; }
    else if (k == 3) { pi = roi - 1 - j; pj = i; }
    const int sy = r.oy + pi, sx = r.ox + pj;
    float vx = 0.f, vy = 0.f;
    if (sy >= r.by0 && sy < r.by1 && sx >= r.bx0 && sx < r.bx1) {
      const size_t idx = base + (size_t)sy * W + sx;
      vx = hu_window(lo[idx], a_min, range);
      vy = hu_window(hi[idx], a_min, range);
    }
    const size_t o = (size_t)blockIdx.x * roi * roi + p;
    x[o] = vx;
    y[o] = vy;
  }
}

// whole-slice windowing (validation / test transform, Mayo.py:158-167): float32 [0,1] images from int16 HU
__global__ void __launch_bounds__(256) window_full_kernel(const short* __restrict__ hu, size_t n, float a_min, float range,
                                                          float* __restrict__ out) {
  mtd_pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = hu_window(hu[i], a_min, range);
}

}  // namespace

extern "C" {

int mtd_hu_foreground_bbox(const short* hu, int S, int H, int W, float a_min, int* bbox, void* stream) {
  MTD_REQUIRE(hu && bbox && S > 0 && H > 0 && W > 0);
  mtd_launch(hu_bbox_kernel, S, 256, 0, (cudaStream_t)stream, hu, H, W, a_min, bbox);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

int mtd_window_crop_patches(const short* lo, const short* hi, int S, int H, int W, const int* patch_tab, int n_patches, int roi,
                            float a_min, float a_max, float* x, float* y, void* stream) {
  MTD_REQUIRE(lo && hi && patch_tab && x && y && S > 0 && H > 0 && W > 0 && n_patches > 0 && roi > 0 && a_max > a_min);
  mtd_launch(window_crop_kernel, n_patches, 256, 0, (cudaStream_t)stream, lo, hi, H, W, reinterpret_cast<const PatchRow*>(patch_tab),
             roi, a_min, (float)((double)a_max - (double)a_min), x, y);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

int mtd_window_slices(const short* hu, long long n, float a_min, float a_max, float* out, void* stream) {
  MTD_REQUIRE(hu && out && n > 0 && a_max > a_min);
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)mtd_sm_count() * 16) blocks = (long long)mtd_sm_count() * 16;
  mtd_launch(window_full_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, hu, (size_t)n, a_min, (float)((double)a_max - (double)a_min), out);
  MTD_CHECK_LAUNCH();
  return MTD_OK;
}

}  // extern "C"
