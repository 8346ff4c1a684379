"""Checkpoint compatibility with the reference's training script.

The reference saves (train.py:276-288) and resumes (train.py:146-159) GAN-style models with the dict
    {'model_state_dict', 'optimizer_D', 'scheduler_D', 'optimizer_G', 'scheduler_G', 'epoch', 'args'}
where `model_state_dict` is `MTD_GAN_Method.state_dict()` (326 entries: Generator.* / Discriminator.*, spectral-norm
layers as bias / weight_orig / weight_u / weight_v) and the optimizers are `torch.optim.AdamW` (optimizers.py:9).
The drop-in modules register the same keys in the same order and `FusedAdamW` keeps torch's per-parameter state keys
('step', 'exp_avg', 'exp_avg_sq'), so checkpoints interchange in both directions; these helpers restate the two
code paths (including the `.module` key fix of :149 for DataParallel-written files) so a resume needs no edits.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

KEYS = ("model_state_dict", "optimizer_D", "scheduler_D", "optimizer_G", "scheduler_G", "epoch", "args")


def checkpoint_dict(model, optimizer_D, scheduler_D, optimizer_G, scheduler_G, epoch: int, args: Any = None) -> Dict[str, Any]:
    """train.py:279-287.  Tensors are moved to the CPU so the file loads anywhere (`map_location='cpu'` at :148)."""
    m = model.module if hasattr(model, "module") else model

    def cpu(o):
        if torch.is_tensor(o):
            return o.detach().cpu()
        if isinstance(o, dict):
            return {k: cpu(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(cpu(v) for v in o)
        return o

    return {"model_state_dict": cpu(m.state_dict()), "optimizer_D": cpu(optimizer_D.state_dict()),
            "scheduler_D": None if scheduler_D is None else scheduler_D.state_dict(), "optimizer_G": cpu(optimizer_G.state_dict()),
            "scheduler_G": None if scheduler_G is None else scheduler_G.state_dict(), "epoch": int(epoch), "args": args}


def save_checkpoint(path: str, model, optimizer_D, scheduler_D, optimizer_G, scheduler_G, epoch: int, args: Any = None):
    torch.save(checkpoint_dict(model, optimizer_D, scheduler_D, optimizer_G, scheduler_G, epoch, args), path)


def fix_optimizer(optimizer):
    """utils.fix_optimizer of the reference (called at train.py:158-159): optimizer state follows its parameter's device.
    FusedAdamW additionally keeps `step` as a device scalar (read by the kernel)."""
    for group in optimizer.param_groups:
        for p in group["params"]:
            st = optimizer.state.get(p)
            if not st:
                continue
            for k, v in st.items():
                if torch.is_tensor(v):
                    st[k] = v.to(device=p.device, dtype=torch.float32 if k == "step" else v.dtype)


def load_checkpoint(path_or_dict, model, optimizer_D=None, scheduler_D=None, optimizer_G=None, scheduler_G=None,
                    strict: bool = True) -> int:
    """train.py:146-159.  Returns the epoch to resume from (`checkpoint['epoch'] + 1`)."""
    ck = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, str) else path_or_dict
    sd = {k.replace(".module", ""): v for k, v in ck["model_state_dict"].items()}         # :149
    model.load_state_dict(sd, strict=strict)
    for opt, key in ((optimizer_D, "optimizer_D"), (optimizer_G, "optimizer_G")):
        if opt is not None and ck.get(key) is not None:
            opt.load_state_dict(ck[key])
            fix_optimizer(opt)
    for sch, key in ((scheduler_D, "scheduler_D"), (scheduler_G, "scheduler_G")):
        if sch is not None and ck.get(key) is not None:
            sch.load_state_dict(ck[key])
    return int(ck.get("epoch", -1)) + 1
