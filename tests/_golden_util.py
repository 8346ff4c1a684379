"""Shared helpers for the golden fixtures (used by tests/golden/make_golden.py and by the tests)."""
import os
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sample_indices(numel: int, k: int = 32) -> torch.Tensor:
    g = torch.Generator().manual_seed(numel % (2 ** 31 - 1) + 17)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def summarize(t: torch.Tensor, k: int = 32) -> dict:
    """Small fingerprint of a big tensor: shape, L2 norm, sum (fp64) and k sampled entries."""
    f = t.detach().double().flatten().cpu()
    idx = sample_indices(f.numel(), k)
    return {"shape": tuple(t.shape), "norm": float(f.norm()), "sum": float(f.sum()),
            "samples": f[idx].float().clone()}


def check_summary(t: torch.Tensor, ref: dict, rtol: float, what: str = ""):
    """Compare a tensor against a stored fingerprint; error is measured relative to the tensor's RMS
    (norm / sqrt(numel)) so near-zero entries do not blow the relative error up."""
    f = t.detach().double().flatten().cpu()
    assert tuple(t.shape) == tuple(ref["shape"]), f"{what}: shape {tuple(t.shape)} vs {ref['shape']}"
    n = f.numel()
    rms = max(ref["norm"] / (n ** 0.5), 1e-30)
    idx = sample_indices(n, len(ref["samples"]))
    err_s = float((f[idx] - ref["samples"].double()).abs().max()) / rms
    err_n = abs(float(f.norm()) - ref["norm"]) / max(ref["norm"], 1e-30)
    assert err_n <= rtol, f"{what}: norm {float(f.norm()):.8e} vs {ref['norm']:.8e} (rel {err_n:.2e} > {rtol})"
    assert err_s <= 10 * rtol, f"{what}: sampled entries differ by {err_s:.2e} x RMS (> {10 * rtol})"


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (norm-wise relative error used for every fp32 parity check)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name), weights_only=False)
