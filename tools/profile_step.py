#!/usr/bin/env python
"""Runs a few MTD-GAN train steps (or 512^2 generator forwards) and brackets ONE steady-state step with
cudaProfilerStart / cudaProfilerStop (process-wide: the backward kernels are launched from autograd's worker thread,
which an NVTX range pushed on the main thread does not cover).  For ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/profile_step.py [--what infer]"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--what", default="train", choices=["train", "infer"])
    args = ap.parse_args()
    from arch.Ours.networks import MTD_GAN_Method
    from module.weight_methods import WeightMethods
    from mtdgan_b200.data import synthetic_pair
    from mtdgan_b200.optim import FusedAdamW
    dev = torch.device("cuda")
    torch.manual_seed(2024)
    random.seed(2024)
    model = MTD_GAN_Method().to(dev).train()
    D, G = model.Discriminator, model.Generator
    if args.what == "infer":
        model.eval()
        x = synthetic_pair(1, 512, seed=4321)[0].to(dev)
        with torch.no_grad():
            for _ in range(args.warm):
                G(x)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            G(x)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return
    opt_D = FusedAdamW([{"params": list(D.parameters())}, {"params": [], "lr": 0.025}], lr=1e-4, weight_decay=5e-4)
    opt_G = FusedAdamW(G.parameters(), lr=1e-4, weight_decay=5e-4)
    wm = WeightMethods("pcgrad", n_tasks=3, device=dev)
    shared, ts, last = list(D.shared_parameters()), list(D.task_specific_parameters()), list(D.last_shared_parameters())
    x, y = (t.to(dev) for t in synthetic_pair(args.batch, 64, seed=1234))

    from mtdgan_b200.graphs import GraphedTrainStep
    graphed = GraphedTrainStep(model, opt_D, opt_G, wm)

    def step():                               # exactly the kernel sequence bench.py captures into its CUDA graph
        graphed.eager_step(x, y)

    for _ in range(args.warm):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
