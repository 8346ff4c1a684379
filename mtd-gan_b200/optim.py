"""Fused multi-tensor AdamW: drop-in for the torch.optim.AdamW the reference builds in
optimizers.py:9 / train.py:122-124 and steps at engine.py:44,52.  Same hyper-parameters, same per-parameter
state keys ('step', 'exp_avg', 'exp_avg_sq') so optimizer state_dicts interchange with torch's; parameters
whose .grad is None are skipped; empty parameter groups are accepted (A12).
One kernel launch per parameter group instead of torch's foreach chain."""
from __future__ import annotations

import torch

from . import _ext
from ._ext import call, ptr, stream
from .weight_methods import _chunk_table


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = 1.0            # multiplies every gradient inside the kernel (1/world of a summed all-reduce)
        self._lr_bufs = {}               # group index -> (pinned host float32[1], device float32[1])

    # The learning rate travels through device memory: step() writes group['lr'] into a pinned host scalar and
    # enqueues its copy to a device scalar the kernel reads.  Captured into a CUDA graph that copy is a memcpy node
    # re-reading the pinned scalar on every replay, so `sync_lr_host()` before a replay is all an lr scheduler
    # (train.py steps scheduler_D / scheduler_G every epoch) needs -- betas / eps / weight decay stay launch constants.
    def _lr_buffers(self, gi, device):
        ent = self._lr_bufs.get(gi)
        if ent is None or ent[1].device != device:
            ent = (torch.zeros(1, dtype=torch.float32).pin_memory(), torch.zeros(1, dtype=torch.float32, device=device))
            self._lr_bufs[gi] = ent
        return ent

    def sync_lr_host(self):
        """Refresh the pinned learning-rate scalars from param_groups (call before replaying a captured step)."""
        for gi, group in enumerate(self.param_groups):
            ent = self._lr_bufs.get(gi)
            if ent is not None:
                ent[0][0] = float(group["lr"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            rows, numels, keep = [], [], []
            device = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW handles fp32 CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    # the step counter lives on the device (like torch's capturable AdamW): the kernel increments it
                    # and derives the bias corrections, so a captured CUDA graph needs no host-side update
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif not st["step"].is_cuda:           # state loaded from a torch.optim.AdamW checkpoint
                    st["step"] = st["step"].to(device=p.device, dtype=torch.float32)
                g = p.grad.contiguous()
                keep.append(g)
                rows.append([p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                             st["step"].data_ptr(), 0, 0])
                numels.append(p.numel())
                device = p.device
            if not rows:
                continue
            seg = _ext.device_table(rows, torch.int64, device)
            chunks, n_chunks = _chunk_table(numels, device)
            lr = float(group["lr"])
            lr_dev = None
            if torch.cuda.is_current_stream_capturing():      # eager launches take the host scalar by value
                lr_host, lr_dev = self._lr_buffers(gi, device)
                lr_host[0] = lr
                lr_dev.copy_(lr_host, non_blocking=True)
            call("mtd_adamw_step", ptr(seg), len(rows), ptr(chunks), n_chunks, lr, ptr(lr_dev), float(b1), float(b2),
                 float(group["eps"]), float(group["weight_decay"]), float(self.grad_scale), stream())
            for p in group["params"]:
                if p.grad is not None:
                    torch.autograd.graph.increment_version(p)      # the kernel wrote p in place (pack caches key on it)
        return loss
