#!/usr/bin/env python
"""Per-shape device time of the conv kernels on the layer shapes of one MTD-GAN train step (B = 20): each shape is
launched `reps` times back to back between two CUDA events (after warm-up), so host launch overhead is amortised.
Prints us per launch, achieved TFLOP/s and the GB/s of compulsory traffic (in + out + weights)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from mtdgan_b200 import ops
    dev = torch.device("cuda")
    B = 20
    shapes = [  # H, C1, C2, N, k, stride, pad, count per step (approx, fwd)
        (64, 32, 0, 32, 3, 1, 1), (64, 64, 0, 64, 3, 1, 1), (32, 64, 0, 128, 3, 1, 1), (32, 128, 0, 128, 3, 1, 1),
        (16, 128, 0, 256, 3, 1, 1), (16, 256, 0, 256, 3, 1, 1), (8, 256, 0, 512, 3, 1, 1), (8, 512, 0, 512, 3, 1, 1),
        (4, 512, 0, 512, 3, 1, 1), (2, 512, 0, 512, 3, 1, 1), (1, 512, 0, 512, 1, 1, 0),
        (2, 512, 512, 512, 3, 1, 1), (4, 512, 512, 512, 3, 1, 1), (8, 512, 512, 256, 3, 1, 1), (16, 256, 256, 128, 3, 1, 1),
        (32, 128, 128, 64, 3, 1, 1), (64, 64, 64, 1, 3, 1, 1), (64, 1, 0, 64, 3, 1, 1), (64, 1, 0, 1, 3, 1, 1),
        (64, 64, 0, 64, 4, 2, 1), (32, 128, 0, 128, 4, 2, 1), (16, 256, 0, 256, 4, 2, 1), (8, 512, 0, 512, 4, 2, 1),
        (4, 512, 0, 512, 4, 2, 1), (2, 512, 0, 512, 4, 2, 1), (1, 512, 0, 2048, 1, 1, 0), (4, 512, 0, 2048, 1, 1, 0),
        (32, 64, 0, 256, 1, 1, 0),
    ]
    reps = 20
    print(f"{'shape (H,C1,C2,N,k,s)':28s} {'fwd us':>9s} {'TF/s':>7s} {'GB/s':>7s} | {'dgrad us':>9s} {'TF/s':>7s} | {'wgrad us':>9s} {'TF/s':>7s}")
    for (H, C1, C2, N, k, s, p) in shapes:
        C = C1 + C2
        x1 = torch.randn(B, H, H, C1, device=dev).requires_grad_(True)
        x2 = torch.randn(B, H, H, C2, device=dev).requires_grad_(True) if C2 else None
        w = (torch.randn(N, C, k, k, device=dev) / (C * k * k) ** 0.5).requires_grad_(True)
        b = torch.zeros(N, device=dev, requires_grad=True)
        cfg = ops.ConvCfg(cin=C, cout=N, kh=k, kw=k, stride=s, pad=p, pre_act=ops.ACT_LEAKY)
        Ho = (H + 2 * p - k) // s + 1
        flop = 2.0 * B * Ho * Ho * N * C * k * k
        byts = 4.0 * (B * H * H * C + B * Ho * Ho * N + N * C * k * k)

        def timeit(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / reps

        with torch.no_grad():
            t_f = timeit(lambda: ops.conv(x1, w, b, cfg, x2=x2))
        y = ops.conv(x1, w, b, cfg, x2=x2)
        g = torch.randn_like(y)
        # dgrad only (weight gradient filtered out) / wgrad only (inputs detached); both include the act_bwd pass
        with ops.wgrad_only_for([b]):
            t_d = timeit(lambda: torch.autograd.grad(y, [x1] + ([x2] if C2 else []), g, retain_graph=True))
        y2 = ops.conv(x1.detach(), w, b, cfg, x2=None if x2 is None else x2.detach())
        t_w = timeit(lambda: torch.autograd.grad(y2, [w], g, retain_graph=True))
        print(f"{str((H, C1, C2, N, k, s)):28s} {t_f:9.1f} {flop / t_f / 1e6:7.1f} {byts / t_f / 1e3:7.0f} | {t_d:9.1f} {flop / t_d / 1e6:7.1f} | "
              f"{t_w:9.1f} {flop / t_w / 1e6:7.1f}", flush=True)


if __name__ == "__main__":
    main()
