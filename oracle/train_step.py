"""One MTD-GAN training iteration on the CPU oracle — the exact sequence of engine.py:40-55 — used as the
`cpu_baseline` / `--impl reference` leg of bench.py and by tests.  TEST / BASELINE INFRASTRUCTURE ONLY."""
from __future__ import annotations

import random
from typing import Dict, Optional, Sequence

import torch

from . import mtdgan_oracle as O


class OracleTrainer:
    """Holds a reference-keyed state dict as trainable leaves plus the two AdamW optimisers
    (train.py:122-127: lr, betas (0.9, 0.999), eps 1e-8, weight decay 5e-4)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], lr: float = 1e-4):
        self.sd = {k: v.detach().clone() for k, v in state_dict.items()}
        for k, v in self.sd.items():
            if not k.endswith(("weight_u", "weight_v")):
                v.requires_grad_(True)
        dn = "Discriminator."
        self.d_shared = [self.sd[dn + n] for n in O.d_shared_names()]
        self.d_ts = [self.sd[dn + n] for n in O.d_task_specific_names()]
        d_params = [v for k, v in self.sd.items() if k.startswith(dn) and v.requires_grad]
        g_params = [v for k, v in self.sd.items() if k.startswith("Generator.")]
        self.d_params, self.g_params = d_params, g_params
        kw = dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4)
        self.opt_D = torch.optim.AdamW(d_params, **kw)
        self.opt_G = torch.optim.AdamW(g_params, **kw)

    def step(self, x, y, dropout_masks: Optional[Sequence[Optional[torch.Tensor]]] = None):
        dm = list(dropout_masks) if dropout_masks is not None else None
        if dm is None:                                       # dropout drawn like nn.Dropout(0.3) would
            B = x.shape[0]
            dm = [torch.nn.functional.dropout(torch.ones(B, 512, device=x.device), 0.3, True) for _ in range(5)]
        # ---- discriminator (engine.py:40-46)
        self.opt_D.zero_grad(set_to_none=True)
        d_losses, d_det = O.d_loss(self.sd, x, y, True, dm[:4])
        grads = [torch.autograd.grad(l, self.d_shared, retain_graph=True) for l in d_losses]
        merged = O.pcgrad_project_lists(grads, "sum")
        for p, g in zip(self.d_shared, merged):
            p.grad = g
        for p, g in zip(self.d_ts, torch.autograd.grad(d_losses.sum(), self.d_ts)):
            p.grad = g
        self.opt_D.step()
        # ---- generator (engine.py:48-55)
        self.opt_G.zero_grad(set_to_none=True)
        for p in self.d_params:
            p.grad = None
        g_loss, g_det = O.g_loss(self.sd, x, y, True, dm[4])
        g_loss.backward()
        self.opt_G.step()
        return d_losses.detach(), d_det, g_loss.detach(), g_det
