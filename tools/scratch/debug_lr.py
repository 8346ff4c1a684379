import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import mtdgan_oracle as O
from arch.Ours.networks import MTD_GAN_Method
from module.weight_methods import WeightMethods
from mtdgan_b200.graphs import GraphedTrainStep
from mtdgan_b200.optim import FusedAdamW
DEV = "cuda"
x, y = (t.to(DEV) for t in O.synthetic_pair(4, 64, seed=9))
for use_graph in (False, True):
    for new_lr in (1e-4, 3e-5):
        torch.manual_seed(2024); random.seed(2024)
        m = MTD_GAN_Method().to(DEV).train()
        m.Discriminator.c_drop.p = 0.0
        D, G = m.Discriminator, m.Generator
        opt_D = FusedAdamW([{"params": list(D.parameters())}, {"params": [], "lr": 0.025}], lr=1e-4, weight_decay=5e-4)
        opt_G = FusedAdamW(G.parameters(), lr=1e-4, weight_decay=5e-4)
        wm = WeightMethods('pcgrad', n_tasks=3, device=torch.device(DEV))
        runner = GraphedTrainStep(m, opt_D, opt_G, wm)
        if use_graph:
            runner.capture(x, y, warmup=1)
        for o in (opt_D, opt_G):
            o.param_groups[0]["lr"] = new_lr
        w0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
        random.seed(3)
        runner(x, y)
        torch.cuda.synchronize()
        sd = m.state_dict()
        out = []
        for k in ("Discriminator.conv12.weight_orig", "Discriminator.conv42.weight_orig", "Discriminator.s_dconv31.weight_orig",
                  "Generator.encoder.3.weight", "Generator.enforce.5.fft_conv.weight"):
            d = (sd[k] - w0[k]).abs()
            out.append("%s mean %.3e max %.3e" % (k.split(".", 1)[1][:22], float(d.mean()), float(d.max())))
        lrs = {gi: float(b[1]) for gi, b in opt_D._lr_bufs.items()}
        print("graph" if use_graph else "eager", "lr", new_lr, "lr_dev", lrs, "step", float(opt_D.state[D.conv12.weight_orig]["step"]), "|", " | ".join(out), flush=True)
