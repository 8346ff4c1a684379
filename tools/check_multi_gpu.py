#!/usr/bin/env python
"""Multi-GPU equality check: R ranks x B patches must equal one process on the concatenated R*B batch —
d_losses averaged over ranks and the post-PCGrad / task-specific / generator gradients (dropout masks sliced from
one global mask).  Run:  torchrun --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 tools/check_multi_gpu.py"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def drop_mask(b, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(b, 512, generator=g) >= 0.3).float() / 0.7


def main():
    from arch.Ours.networks import MTD_GAN_Method
    from module.weight_methods import WeightMethods
    from mtdgan_b200 import distributed as mdist, networks as NW
    from mtdgan_b200.data import synthetic_pair
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = 4
    xg, yg = synthetic_pair(B * world, 64, seed=77)
    masks_global = [drop_mask(B * world, 300 + i) for i in range(5)]

    def run(x, y, masks, distributed):
        torch.manual_seed(2024)
        random.seed(2024)
        m = MTD_GAN_Method().to(dev).train()
        D, G = m.Discriminator, m.Generator
        q = [t.to(dev) for t in masks]
        NW.set_dropout_mask_provider(lambda b, n, d: q.pop(0))
        try:
            wm = WeightMethods("pcgrad", n_tasks=3, device=dev)
            d_losses, _ = m.d_loss(x.to(dev), y.to(dev))
            if not distributed:
                # single-process path must not see the process group
                saved = mdist.active
                mdist.active = lambda: False
            try:
                wm.backward(losses=d_losses, shared_parameters=list(D.shared_parameters()),
                            task_specific_parameters=list(D.task_specific_parameters()),
                            last_shared_parameters=list(D.last_shared_parameters()))
            finally:
                if not distributed:
                    mdist.active = saved
            g_loss, _ = m.g_loss(x.to(dev), y.to(dev))
            g_loss.backward()
            if distributed:
                mdist.allreduce_mean_grads(list(G.parameters()))
        finally:
            NW.set_dropout_mask_provider(None)
        grads = {k: p.grad.detach().clone() for k, p in list(D.named_parameters()) + [("G." + k, p) for k, p in G.named_parameters()]
                 if p.grad is not None}
        return d_losses.detach().clone(), g_loss.detach().clone(), grads

    mdist.init()
    sl = slice(rank * B, (rank + 1) * B)
    dl, gl, grads = run(xg[sl], yg[sl], [m[sl] for m in masks_global], True)
    dl_avg = dl.clone()
    dist.all_reduce(dl_avg)
    dl_avg /= world
    gl_avg = gl.clone()
    dist.all_reduce(gl_avg)
    gl_avg /= world
    ok = True
    if rank == 0:
        dl1, gl1, grads1 = run(xg, yg, masks_global, False)
        print("d_losses   ranks-avg", dl_avg.tolist(), "single", dl1.tolist())
        print("g_loss     ranks-avg", float(gl_avg), "single", float(gl1))
        worst, errs = ("", 0.0), []
        for k, g1 in grads1.items():
            e = float((grads[k] - g1).abs().max() / g1.abs().max().clamp_min(1e-30))
            errs.append(e)
            if e > worst[1]:
                worst = (k, e)
        errs.sort()
        print(f"gradients: {len(errs)} tensors, median rel err {errs[len(errs) // 2]:.3e}, worst {worst[1]:.3e} ({worst[0]})")
        ok = (torch.allclose(dl_avg[:2], dl1[:2], rtol=1e-4) and abs(float(gl_avg) - float(gl1)) <= 1e-4 * abs(float(gl1))
              and errs[len(errs) // 2] <= 1e-3)
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
