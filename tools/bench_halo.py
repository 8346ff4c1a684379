#!/usr/bin/env python
"""Microbenchmark of the discriminator's 3x3 / stride-1 layers at the B = 20 training geometry: streamed-weight halo-tile
kernel vs tap-streaming kernel (forward; CUDA-graph replay of 10 back-to-back launches, L2-warm like the real step)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

SHAPES = [  # B, H, C1, C2, N
    (20, 64, 64, 0, 64), (20, 64, 64, 64, 64), (20, 32, 128, 0, 128), (20, 32, 64, 0, 128), (20, 32, 128, 128, 128),
    (20, 16, 256, 0, 256), (20, 16, 128, 0, 256), (20, 16, 256, 256, 256), (60, 64, 64, 0, 64), (60, 32, 128, 0, 128),
    (60, 16, 256, 0, 256),
]


def main():
    from mtdgan_b200 import _ext, ops
    lib = _ext.load()
    dev = torch.device("cuda")
    for (B, H, C1, C2, N) in SHAPES:
        C = C1 + C2
        x1 = torch.randn(B, H, H, C1, device=dev)
        x2 = torch.randn(B, H, H, C2, device=dev) if C2 else None
        w = torch.randn(N, C, 3, 3, device=dev) / (3.0 * C ** 0.5)
        b = torch.zeros(N, device=dev)
        cfg = ops.ConvCfg(cin=C, cout=N, kh=3, kw=3, stride=1, pad=1, pre_act=ops.ACT_LEAKY)
        flop = 2.0 * B * H * H * C * N * 9
        for passes in (3, 1):
            ops.set_conv_mode("auto", passes)
            line = f"B={B:3d} {H:2d}x{H:<2d} C={C1}+{C2} N={N} passes={passes}:"
            for halo in (1, 0):
                lib.mtd_tc_set_halo(halo)
                with torch.no_grad():
                    for _ in range(3):
                        ops.conv(x1, w, b, cfg, x2=x2)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for _ in range(10):
                            y = ops.conv(x1, w, b, cfg, x2=x2)
                    g.replay(); torch.cuda.synchronize()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
                    us = s.elapsed_time(e) * 1e3 / 10
                line += f"  {'halo' if halo else 'taps'} {us:7.1f} us {flop / us / 1e6:6.1f} TF/s"
            print(line, flush=True)
        lib.mtd_tc_set_halo(1)
        ops.set_conv_mode("auto", 3)


if __name__ == "__main__":
    main()
