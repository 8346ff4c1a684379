#!/usr/bin/env python
"""Tile / split-K sweep of the tcgen05 conv kernels over the layer shapes of one MTD-GAN train step (B = 20).

Every (kernel version, Cout tile, stream-K piece length) candidate is forced through `mtd_tc_set_tuning`, captured as a CUDA graph
of `reps` launches (so host launch cost is out of the picture) that cycle through enough weight copies to defeat
the L2 (in the real step the discriminator's 250 MB of weights never stay resident), and timed with CUDA events.
Prints, per shape, the measured time of every candidate, the best one, and what the built-in cost model picks —
the data the cost model in conv_tc.cu (`choose_tiling*`) is calibrated on.

    python tools/tune_tc.py [--batch 20] [--passes 3] [--quick]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", type=int, default=-1, help="index of a single shape to sweep")
    ap.add_argument("--auto-only", action="store_true", help="only time the cost model's choice (for profiling)")
    ap.add_argument("--versions", default="1,2", help="kernel generations to sweep")
    args = ap.parse_args()
    from mtdgan_b200 import ops, _ext
    from mtdgan_b200._ext import call, fptr
    dev = torch.device("cuda")
    B = args.batch
    # (H, C1, C2, N, k): stride-1 layers of G (64x64 patches) and D; dgrad shapes are the same GEMMs with C and N swapped
    shapes = [
        (64, 32, 0, 32, 3), (64, 64, 0, 64, 1), (64, 64, 0, 64, 3), (64, 128, 0, 64, 3), (32, 64, 0, 128, 3),
        (32, 128, 0, 128, 3), (32, 128, 128, 64, 3), (32, 256, 0, 128, 3),
        (16, 128, 0, 256, 3), (16, 256, 0, 256, 3), (16, 256, 256, 128, 3), (16, 512, 0, 256, 3),
        (8, 256, 0, 512, 3), (8, 512, 0, 512, 3), (8, 512, 512, 256, 3), (8, 1024, 0, 512, 3),
        (4, 512, 0, 512, 3), (4, 512, 512, 512, 3), (2, 512, 0, 512, 3), (2, 512, 512, 512, 3),
        (4, 512, 0, 2048, 1), (32, 64, 0, 256, 1), (2, 512, 0, 512, 1),
    ]
    if args.quick:
        shapes = shapes[::3]
    if args.only >= 0:
        shapes = [shapes[args.only]]
    lib = _ext.load()
    versions = [int(v) for v in args.versions.split(",")]
    for (H, C1, C2, N, k) in shapes:
        if lib.mtd_conv_fwd_tc_supported(B, H, H, C1, C2, N, k, k, 1, k // 2) != 1:
            print(f"{(H, C1, C2, N, k)}: not a tensor-core shape", flush=True)
            continue
        C = C1 + C2
        wbytes = 4 * N * C * k * k * (2 if args.passes == 3 else 1)
        ncopy = max(2, min(16, (160 << 20) // wbytes + 1))
        cfg = ops.ConvCfg(cin=C, cout=N, kh=k, kw=k, stride=1, pad=k // 2, pre_act=ops.ACT_LEAKY)
        kind = "fwd_tf32x3" if args.passes == 3 else "fwd_tf32"
        wps = []
        for _ in range(ncopy):
            w = torch.randn(N, C, k, k, device=dev) / (C * k * k) ** 0.5
            wps.append(ops._packed(w, kind, cfg).clone())
        x1 = torch.randn(B, H, H, C1, device=dev)
        x2 = torch.randn(B, H, H, C2, device=dev) if C2 else None
        bias = torch.zeros(N, device=dev)
        y = torch.empty(B, H, H, N, device=dev)
        ws = torch.empty(8 << 20, device=dev)
        reps = ncopy * (2 if ncopy >= 8 else 8)
        flop = 2.0 * B * H * H * N * C * k * k

        def launch(i, st):
            call("mtd_conv_fwd_tc", fptr(x1), fptr(x2), fptr(wps[i % ncopy]), fptr(bias), None, 0, fptr(y), None, None, None,
                 B, H, H, C1, C2, N, k, k, 1, k // 2, ops.ACT_LEAKY, 0, 0.2, args.passes, fptr(ws), ws.numel(), st)

        def measure():
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for i in range(2):
                    launch(i, s.cuda_stream)
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for i in range(reps):
                        launch(i, torch.cuda.current_stream().cuda_stream)
                g.replay()
                s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(3):
                    g.replay()
                e1.record(s)
                s.synchronize()
            return e0.elapsed_time(e1) * 1e3 / (3 * reps)

        m = B * H * H
        kiters = k * k * C // 32
        results = {}
        for ver in versions:
            ops.set_tc_version(ver)
            call("mtd_tc_set_tuning", 0, 0)
            results[(ver, -1, 0)] = measure()
            if args.auto_only:
                print(f"v{ver} auto: {results[(ver, -1, 0)]:.1f} us", flush=True)
                continue
            for bn in (128, 64, 32):
                if N % bn:
                    continue
                mn = ((m + 127) // 128) * (N // bn)
                rem = mn % 148
                call("mtd_tc_set_tuning", bn, -1)             # whole tiles only
                results[(ver, bn, 0)] = measure()
                if rem == 0 or ver == 2:
                    continue
                per_min = -(-rem * kiters // 148)
                seen = set()
                for mult in (1.0, 1.25, 1.5, 2.0, 3.0, 4.0, 6.0, 8.0):
                    per = max(2, int(per_min * mult + 0.5))
                    if per in seen or (per >= kiters and mn < 148):
                        continue
                    seen.add(per)
                    call("mtd_tc_set_tuning", bn, per)        # stream-K wave with `per` k-steps per CTA
                    results[(ver, bn, per)] = measure()
        call("mtd_tc_set_tuning", 0, 0)
        if args.auto_only:
            continue
        best = min((v, kk) for kk, v in results.items() if kk[1] > 0)
        line = f"shape H={H} C={C1}+{C2} N={N} k={k} (M={m}, kiters={kiters}, {ncopy} weight copies)  {flop / 1e9:.2f} GFLOP"
        print(line)
        for ver in versions:
            auto = results[(ver, -1, 0)]
            vbest = min((v, kk) for kk, v in results.items() if kk[0] == ver and kk[1] > 0)
            print(f"  v{ver}: auto {auto:7.1f} us ({flop / auto / 1e6:6.1f} TF/s)   best {vbest[0]:7.1f} us @ bn={vbest[1][1]} per={vbest[1][2]}"
                  f"   auto/best = {auto / vbest[0]:.2f}")
            for bn in (128, 64, 32):
                row = [(kk[2], v) for kk, v in sorted(results.items()) if kk[0] == ver and kk[1] == bn]
                if row:
                    print(f"      bn={bn:3d}: " + "  ".join((f"per{ks}={v:.1f}" if ks else f"whole={v:.1f}") for ks, v in row))
        print(f"  overall best: {best[0]:.1f} us  v{best[1][0]} bn={best[1][1]} per={best[1][2]}", flush=True)
    ops.set_tc_version(1)


if __name__ == "__main__":
    main()
