"""CPU restatement of the reference's `window_patch` / `window` transform chains — TEST INFRASTRUCTURE ONLY.

Reference call site: create_datasets/Mayo.py:117-136 (train, `window_patch`) and :158-167 (valid / test, `window`).
The transforms themselves live in a third-party dependency that is ABSENT from /root/reference and from this image:
**monai==1.3.2** (requirements.txt:14; pydicom==2.4.4 for the DICOM decode, :15).  Their published algorithms are
restated here in numpy / torch-CPU, each function naming the MONAI class it follows:

  ScaleIntensityRanged(a_min=-160, a_max=240, b_min=0, b_max=1, clip=True)  -> scale_intensity_range
  CropForegroundd(source_key="n_100", select_fn=x > 0)                      -> foreground_bbox
  SpatialPadd(spatial_size=(64, 64))  (method="symmetric", constant 0)      -> inside crop_window
  RandSpatialCropSamplesd(roi_size=(64, 64), num_samples=8, random_center=True, random_size=False)
                                                                            -> random_crop_origin
  RandRotate90d(prob=0.1, spatial_axes=[0, 1]), RandFlipd(prob=0.1, spatial_axis=[0, 1])   -> draw_decisions
  RandRotated(prob=0.1, +-15 degrees, bilinear)                              -> NOT restated (out of scope)

Parity status: **unpinned** against MONAI itself (it cannot be imported here and the reference ships no golden
vectors for its data pipeline); the arithmetic of scale_intensity_range is pinned against torch-CPU float32 ops in
tests/test_data_edge.py.  What the GPU path is held to, bit for bit, is THIS restatement with the same numpy
RandomState: one RandomState drives all random transforms in call order (MONAI derives one RandomState per transform
from the Compose seed; that derivation is not reproducible without MONAI).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def scale_intensity_range(hu: np.ndarray, a_min=-160.0, a_max=240.0, b_min=0.0, b_max=1.0, clip=True) -> np.ndarray:
    """monai.transforms.ScaleIntensityRange.__call__ (1.3.2): tensor arithmetic with Python-float scalars on an int16
    image promotes to float32."""
    img = torch.as_tensor(np.ascontiguousarray(hu))
    img = (img - a_min) / (a_max - a_min)
    img = img * (b_max - b_min) + b_min
    if clip:
        img = torch.clamp(img, b_min, b_max)
    return img.to(torch.float32).numpy()


def foreground_bbox(win_hi: np.ndarray) -> Tuple[int, int, int, int]:
    """monai.transforms.utils.generate_spatial_bounding_box with select_fn = x > 0, margin 0 on a (H, W) image:
    (y0, y1, x0, x1) half-open; all zeros when nothing is selected."""
    fg = win_hi > 0
    if not fg.any():
        return 0, 0, 0, 0
    ys, xs = np.where(fg.any(axis=1))[0], np.where(fg.any(axis=0))[0]
    return int(ys[0]), int(ys[-1]) + 1, int(xs[0]), int(xs[-1]) + 1


def random_crop_origin(size: Tuple[int, int], roi: int, rng: np.random.RandomState) -> Tuple[int, int]:
    """monai.data.utils.get_random_patch: per dimension `rand_int(ms - ps + 1) if ms > ps else 0` (no draw otherwise)."""
    return tuple(int(rng.randint(ms - roi + 1)) if ms > roi else 0 for ms in size)


def draw_decisions(size: Tuple[int, int], roi: int, num_samples: int, rng: np.random.RandomState, prob_rot=0.1,
                   prob_flip=0.1) -> List[tuple]:
    """Random decisions of one slice in Compose order (each transform is mapped over the list of samples before the next
    one runs): num_samples crop origins; then per sample RandRotate90d.randomize (`_rand_k = R.randint(max_k) + 1`, then
    `_do_transform = R.rand() < prob`); then per sample RandFlipd.randomize (`R.rand() < prob`).
    Returns [(oy, ox, k or 0, flip 0/1)]."""
    origins = [random_crop_origin(size, roi, rng) for _ in range(num_samples)]
    rots = []
    for _ in range(num_samples):
        k = int(rng.randint(3)) + 1
        rots.append(k if rng.rand() < prob_rot else 0)
    flips = [int(rng.rand() < prob_flip) for _ in range(num_samples)]
    return [(oy, ox, k, f) for (oy, ox), k, f in zip(origins, rots, flips)]


def window_patch_pipeline(hu_lo: np.ndarray, hu_hi: np.ndarray, rng: np.random.RandomState, roi=64, num_samples=8,
                          a_min=-160.0, a_max=240.0) -> Tuple[np.ndarray, np.ndarray, List[tuple]]:
    """One (low-dose, full-dose) slice pair through Mayo.py:117-136 (minus RandRotated).
    Returns x, y of shape (num_samples, 1, roi, roi) float32 and the per-sample decisions (oy, ox, k, flip)."""
    lo, hi = scale_intensity_range(hu_lo, a_min, a_max), scale_intensity_range(hu_hi, a_min, a_max)
    y0, y1, x0, x1 = foreground_bbox(hi)
    lo, hi = lo[y0:y1, x0:x1], hi[y0:y1, x0:x1]                      # CropForegroundd (both keys, box from n_100)
    pads = []
    for n in lo.shape:                                               # SpatialPadd, symmetric
        w = max(roi - n, 0)
        pads.append((w // 2, w - w // 2))
    lo, hi = np.pad(lo, pads), np.pad(hi, pads)
    xs, ys = [], []
    dec = draw_decisions(lo.shape, roi, num_samples, rng)
    for oy, ox, k, flip in dec:                                      # RandSpatialCropSamplesd
        pl, ph = lo[oy:oy + roi, ox:ox + roi], hi[oy:oy + roi, ox:ox + roi]
        if k:
            pl, ph = np.rot90(pl, k), np.rot90(ph, k)                # RandRotate90d, spatial_axes (0, 1)
        if flip:
            pl, ph = np.flip(pl, (0, 1)), np.flip(ph, (0, 1))        # RandFlipd, spatial_axis [0, 1]
        xs.append(np.ascontiguousarray(pl)[None])
        ys.append(np.ascontiguousarray(ph)[None])
    return np.stack(xs).astype(np.float32), np.stack(ys).astype(np.float32), dec
