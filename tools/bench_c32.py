#!/usr/bin/env python
"""Microbenchmark of the 32 -> 32 channel 3x3 layer (the generator's spatial conv) at the inference and training
geometries: halo-tile kernel vs general kernel, 3xTF32 vs plain TF32, CUDA-graph replay of 20 back-to-back launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from mtdgan_b200 import _ext, ops
    lib = _ext.load()
    dev = torch.device("cuda")
    for (B, H) in ((1, 512), (4, 512), (20, 64), (40, 64)):
        x = torch.randn(B, H, H, 32, device=dev)
        w = torch.randn(32, 32, 3, 3, device=dev) / 17.0
        b = torch.zeros(32, device=dev)
        skip = torch.randn(B, H, H, 32, device=dev)
        cfg = ops.ConvCfg(cin=32, cout=32, kh=3, kw=3, stride=1, pad=1, pre_act=ops.ACT_RELU)
        flop = 2.0 * B * H * H * 32 * 32 * 9
        for passes in (3, 1):
            ops.set_conv_mode("auto", passes)
            for c32 in (1, 0):
                lib.mtd_tc_set_c32(c32)
                with torch.no_grad():
                    for _ in range(3):
                        ops.conv(x, w, b, cfg, add1=skip)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for _ in range(20):
                            y = ops.conv(x, w, b, cfg, add1=skip)
                    g.replay(); torch.cuda.synchronize()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
                    us = s.elapsed_time(e) * 1e3 / 20
                print(f"B={B:3d} {H}x{H} passes={passes} {'halo' if c32 else 'general':8s} {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s  "
                      f"{(3 * x.numel() * 4) / us / 1e3:7.0f} GB/s (in+skip+out)", flush=True)
        lib.mtd_tc_set_c32(1)
        ops.set_conv_mode("auto", 3)


if __name__ == "__main__":
    main()
