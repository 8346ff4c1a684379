"""Synthetic low-dose / full-dose patch pairs for benchmarks and smoke tests: the real Mayo
data are private (README.md:123-125 of the reference).  y = clamp(1.6*rand - 0.3, 0, 1) has plateaus at
exactly 0 and 1 like HU windowing with clip=True (create_datasets/Mayo.py:120); x = clamp(y + 0.1*randn, 0, 1),
so ~19 % of pixels have x == y exactly and the NDS mask is non-trivial."""
import torch


def synthetic_pair(batch: int, size: int, seed: int = 1234, rank: int = 0, pin: bool = False):
    g = torch.Generator().manual_seed(seed + rank)
    base = torch.rand(batch, 1, size, size, generator=g)
    y = torch.clamp(1.6 * base - 0.3, 0, 1)
    x = torch.clamp(y + 0.1 * torch.randn(batch, 1, size, size, generator=g), 0, 1)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x, y
