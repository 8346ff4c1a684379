"""Data edge of the hot path.

* `synthetic_pair`: synthetic low-dose / full-dose patch pairs for benchmarks and smoke tests: the real
  Mayo data are private (README.md:123-125 of the reference).  y = clamp(1.6*rand - 0.3, 0, 1) has plateaus at exactly 0
  and 1 like HU windowing with clip=True (create_datasets/Mayo.py:120); x = clamp(y + 0.1*randn, 0, 1), so ~19 % of
  pixels have x == y exactly and the NDS mask is non-trivial.
* `WindowPatchSampler`: the reference's `window_patch` training transform chain (create_datasets/Mayo.py:117-136:
  HU window [-160, 240] -> [0, 1] with clip, foreground crop on the full-dose slice, pad to 64 x 64, 8 random 64 x 64
  crops per slice, random rot90 / flip) on int16 HU slices resident on the GPU, replacing the MONAI CPU workers that
  become the bottleneck once a train step takes ~40 ms.  The DICOM decode (pydicom) stays on the host and the
  +-15 degree RandRotated (prob 0.1) is not reproduced.
* `window_slices`: the validation / test transform (Mayo.py:158-167).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _ext
from ._ext import call, fptr, ptr, stream


def synthetic_pair(batch: int, size: int, seed: int = 1234, rank: int = 0, pin: bool = False):
    g = torch.Generator().manual_seed(seed + rank)
    base = torch.rand(batch, 1, size, size, generator=g)
    y = torch.clamp(1.6 * base - 0.3, 0, 1)
    x = torch.clamp(y + 0.1 * torch.randn(batch, 1, size, size, generator=g), 0, 1)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


def _check_hu(t: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _ext.MtdError(f"{what}: expected a CUDA tensor — the B200 data edge has no CPU fallback")
    if t.dtype != torch.int16 or t.dim() != 3:
        raise _ext.MtdError(f"{what}: expected int16 HU slices of shape (S, H, W), got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def window_slices(hu: torch.Tensor, a_min: float = -160.0, a_max: float = 240.0) -> torch.Tensor:
    """ScaleIntensityRanged(a_min, a_max, 0, 1, clip=True) + AddChanneld: (S, H, W) int16 -> (S, 1, H, W) float32."""
    hu = _check_hu(hu, "window_slices")
    out = torch.empty((hu.shape[0], 1, hu.shape[1], hu.shape[2]), dtype=torch.float32, device=hu.device)
    call("mtd_window_slices", ptr(hu), hu.numel(), float(a_min), float(a_max), fptr(out), stream())
    return out


def draw_decisions(size: Tuple[int, int], roi: int, num_samples: int, rng: np.random.RandomState, prob_rot=0.1,
                   prob_flip=0.1) -> List[tuple]:
    """Random decisions of one slice in the reference's Compose order (Mayo.py:126-131): `num_samples` crop origins
    (per dimension one `randint(size - roi + 1)`, no draw when size == roi); then per sample the rot90 count
    (`randint(3) + 1`) and its coin (`rand() < 0.1`); then per sample the flip coin."""
    origins = [tuple(int(rng.randint(ms - roi + 1)) if ms > roi else 0 for ms in size) for _ in range(num_samples)]
    rots = []
    for _ in range(num_samples):
        k = int(rng.randint(3)) + 1
        rots.append(k if rng.rand() < prob_rot else 0)
    flips = [int(rng.rand() < prob_flip) for _ in range(num_samples)]
    return [(oy, ox, k, f) for (oy, ox), k, f in zip(origins, rots, flips)]


class WindowPatchSampler:
    """(hu_low, hu_high): int16 (S, H, W) CUDA tensors of paired slices -> (x, y): float32 (S * num_samples, 1, roi, roi),
    sample-major per slice like `list_data_collate` of the reference's loader (Mayo.py:187)."""

    def __init__(self, roi: int = 64, num_samples: int = 8, a_min: float = -160.0, a_max: float = 240.0,
                 prob_rot90: float = 0.1, prob_flip: float = 0.1, seed: Optional[int] = None):
        self.roi, self.n, self.a_min, self.a_max = int(roi), int(num_samples), float(a_min), float(a_max)
        self.prob_rot, self.prob_flip = prob_rot90, prob_flip
        self.rng = np.random.RandomState(seed)
        self.last_decisions: List[List[tuple]] = []

    def __call__(self, hu_low: torch.Tensor, hu_high: torch.Tensor):
        lo, hi = _check_hu(hu_low, "WindowPatchSampler"), _check_hu(hu_high, "WindowPatchSampler")
        if lo.shape != hi.shape:
            raise _ext.MtdError(f"low-dose / full-dose slice stacks differ in shape: {tuple(lo.shape)} vs {tuple(hi.shape)}")
        S, H, W = lo.shape
        bbox = torch.empty((S, 4), dtype=torch.int32, device=lo.device)
        call("mtd_hu_foreground_bbox", ptr(hi), S, H, W, self.a_min, ptr(bbox), stream())
        boxes = bbox.cpu().tolist()            # 16 bytes per slice: the crop ranges depend on the boxes (host RNG)
        rows, self.last_decisions = [], []
        for s, (y0, y1, x0, x1) in enumerate(boxes):
            size, pad = [], []
            for n in (y1 - y0, x1 - x0):       # SpatialPadd(spatial_size=roi), symmetric
                w = max(self.roi - n, 0)
                pad.append(w // 2)
                size.append(n + w)
            dec = draw_decisions(tuple(size), self.roi, self.n, self.rng, self.prob_rot, self.prob_flip)
            self.last_decisions.append(dec)
            for oy, ox, k, f in dec:           # crop pixel (0, 0) in slice coordinates: box origin - padding + offset
                rows.append([s, y0 - pad[0] + oy, x0 - pad[1] + ox, y0, y1, x0, x1, k | (f << 2)])
        tab = _ext.device_table(rows, torch.int32, lo.device)
        x = torch.empty((len(rows), 1, self.roi, self.roi), dtype=torch.float32, device=lo.device)
        y = torch.empty_like(x)
        call("mtd_window_crop_patches", ptr(lo), ptr(hi), S, H, W, ptr(tab), len(rows), self.roi, self.a_min, self.a_max,
             fptr(x), fptr(y), stream())
        return x, y
