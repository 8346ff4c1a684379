"""Whole-step CUDA-graph capture of the MTD-GAN training iteration.

In eager mode one step issues ~5 000 kernel launches from Python and is bound by the host (~25 us per launch);
captured once, the same kernels replay back-to-back from a single `cudaGraphLaunch`.  Everything in the step is
capture-safe by construction: no host synchronisation (the PCGrad sign tests run on the device), every host-built
table goes through pinned staging buffers, AdamW's step counters live on the device, dropout uses torch's
graph-registered Philox generator, and all scratch memory comes from the graph's private pool.

Host work that must still happen per step is one call: `PCGrad.draw_orders_host()` consumes Python's `random`
exactly like the reference's three `random.shuffle` calls and refreshes the pinned buffer the captured copy node
reads.
"""
from __future__ import annotations

import random

import torch

from . import _ext, ops


class GraphedTrainStep:
    """engine.train_MTD_GAN_Ours' per-batch body (engine.py:40-55) as one replayable CUDA graph.

    model: MTD_GAN_Method; opt_D / opt_G: mtdgan_b200.optim.FusedAdamW; wm: WeightMethods('pcgrad').
    Call with (x, y) of the captured shape; returns (d_losses[3], d_details, g_loss, g_details) whose tensors are
    overwritten by the next call.
    """

    def __init__(self, model, opt_D, opt_G, wm, post_g_backward=None):
        self.model, self.opt_D, self.opt_G, self.wm = model, opt_D, opt_G, wm
        D = model.Discriminator
        self._shared = list(D.shared_parameters())
        self._ts = list(D.task_specific_parameters())
        self._last = list(D.last_shared_parameters())
        self._post_g_backward = post_g_backward      # e.g. the generator-gradient all-reduce at N > 1
        self._d_params = list(D.parameters())
        # Re-packing is restricted to THIS model's parameters: the pack registry is process-wide, and a captured step must
        # never contain writes into the pack buffers of another (possibly soon-to-be-freed) model -- a replay would then
        # write into memory that has been handed to someone else.
        self._all_params = list(model.parameters())
        self.graph = None
        self.x = self.y = None
        self.out = None

    def eager_step(self, x, y):
        m, D, G = self.model, self.model.Discriminator, self.model.Generator
        ops.repack_stale(self._all_params)    # all weights changed in the previous step: one batched re-pack
        self.opt_D.zero_grad(); D.zero_grad()
        d_losses, d_det = m.d_loss(x, y)
        self.wm.backward(losses=d_losses, shared_parameters=self._shared, task_specific_parameters=self._ts,
                         last_shared_parameters=self._last)
        self.opt_D.step()
        ops.repack_stale(self._d_params)      # g_loss runs the updated discriminator
        self.opt_G.zero_grad(); G.zero_grad()
        g_loss, g_det = m.g_loss(x, y)
        g_loss.backward()
        if self._post_g_backward is not None:
            self._post_g_backward()
        self.opt_G.step()
        return d_losses.detach(), d_det, g_loss.detach(), g_det

    # ---- training state that the warm-up steps of a capture must not disturb -----------------------------------
    def _snapshot(self):
        opt = []
        for o in (self.opt_D, self.opt_G):
            opt.append({p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                        for p, st in o.state.items()})
        return {"model": {k: v.detach().clone() for k, v in self.model.state_dict().items()}, "opt": opt,
                "py": random.getstate(), "cpu": torch.get_rng_state(), "cuda": torch.cuda.get_rng_state()}

    @torch.no_grad()
    def _restore(self, snap):
        sd = self.model.state_dict()
        for k, v in snap["model"].items():
            sd[k].copy_(v)                         # in place: parameter storage (and the graph's pointers) stay put
        for p in self.model.parameters():
            torch.autograd.graph.increment_version(p)
        for o, saved in zip((self.opt_D, self.opt_G), snap["opt"]):
            for p, st in o.state.items():
                old = saved.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):         # state created by the warm-up steps goes back to its initial zeros
                        v.zero_() if old is None else v.copy_(old[k])
        random.setstate(snap["py"])
        torch.set_rng_state(snap["cpu"])
        torch.cuda.set_rng_state(snap["cuda"])

    def capture(self, x, y, warmup: int = 3, preserve_state: bool = True):
        """Warm up on a side stream (lazy one-time work: table builds, workspace allocation), then record one step.
        preserve_state: weights, spectral-norm buffers, optimizer moments / step counters, Python `random` and the torch
        RNGs are restored afterwards, so the first replay is training step 1 exactly as an eager run would execute it
        (the warm-up steps are real optimizer steps on the capture batch and would otherwise shift the trajectory)."""
        self.x, self.y = x.clone(), y.clone()
        snap = self._snapshot() if preserve_state else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.eager_step(self.x, self.y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.clear_pack_cache()            # every weight-packing kernel must be part of the captured step
        self.graph = torch.cuda.CUDAGraph()
        _ext.start_trace()
        try:
            with torch.cuda.graph(self.graph):
                self.out = self.eager_step(self.x, self.y)
        finally:
            self.trace = _ext.stop_trace()        # (entry point, args) of every library call in the captured step
        if snap is not None:
            self._restore(snap)
            torch.cuda.synchronize()
        return self

    def family_graph(self, names):
        """A CUDA graph that re-issues, in order, the captured step's calls to the given entry points on the captured
        step's own buffers (graph-private pool).  For timing one kernel family in isolation; only for stateless
        families (convolutions, FFT passes) -- replaying optimizer or power-iteration calls would change the model."""
        lib = _ext.load()
        calls = [(n, a) for n, a in self.trace if n in names]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            for n, a in calls:
                rc = getattr(lib, n)(*a[:-1], st)         # the stream is the last argument of every entry point
                if rc != 0:
                    raise _ext.MtdError(f"{n}: status {rc} while building a family graph")
        return g, len(calls)

    def __call__(self, x, y):
        if self.graph is None:
            return self.eager_step(x, y)
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.wm.method.draw_orders_host()          # same `random` consumption as the reference's shuffles
        self.opt_D.sync_lr_host()                  # lr schedulers: the captured copy nodes re-read the pinned scalars
        self.opt_G.sync_lr_host()
        self.graph.replay()
        return self.out
