import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import mtdgan_oracle as O
from arch.Ours.networks import MTD_GAN_Method
from module.weight_methods import WeightMethods
from mtdgan_b200.graphs import GraphedTrainStep
from mtdgan_b200.optim import FusedAdamW
from mtdgan_b200 import ops
DEV = "cuda"
warm = int(os.environ.get("DBG_WARM", "1"))
preserve = os.environ.get("DBG_PRESERVE", "1") == "1"
norepack = os.environ.get("DBG_NOREPACK", "0") == "1"
if norepack:
    ops.repack_stale = lambda params=None: 0
x, y = (t.to(DEV) for t in O.synthetic_pair(4, 64, seed=9))
for arm in ("eager", "graph", "graph", "eager", "graph"):
    torch.manual_seed(2024); random.seed(2024)
    m = MTD_GAN_Method().to(DEV).train()
    m.Discriminator.c_drop.p = 0.0
    D, G = m.Discriminator, m.Generator
    opt_D = FusedAdamW([{"params": list(D.parameters())}, {"params": [], "lr": 0.025}], lr=1e-4, weight_decay=5e-4)
    opt_G = FusedAdamW(G.parameters(), lr=1e-4, weight_decay=5e-4)
    wm = WeightMethods('pcgrad', n_tasks=3, device=torch.device(DEV))
    runner = GraphedTrainStep(m, opt_D, opt_G, wm)
    if arm == "graph":
        runner.capture(x, y, warmup=warm, preserve_state=preserve)
    w0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    random.seed(3)
    runner(x, y)
    torch.cuda.synchronize()
    bad = []
    for k, v in m.state_dict().items():
        if k.endswith(("weight_u", "weight_v")):
            continue
        d = float((v - w0[k]).abs().max())
        if not (d <= 2.5e-4):
            bad.append((k, round(d, 4), v.numel()))
    print(f"{arm} warm={warm} preserve={preserve} norepack={norepack} PDL={os.environ.get('MTD_PDL','1')} WGS={os.environ.get('MTDGAN_WGRAD_STREAM','1')}: "
          f"{len(bad)} bad", bad[:5], flush=True)
