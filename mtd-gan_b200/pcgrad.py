"""Optimizer-wrapping PCGrad (module/pcgrad.py:13-141) on the B200 path.

Same surface (`PCGrad(optimizer, reduction='mean')`, `.optimizer`, `.zero_grad()`, `.step()`,
`.pc_backward(objectives)`); the flatten / project / unflatten of the reference (:50-70, :85-115) is one
Gram + solve + combine sequence on the device over the per-parameter gradients (nothing is concatenated).
Reference behaviour kept: entries every task has a gradient for are AVERAGED whatever `reduction` says
(the 'sum' branch at :63-65 is unreachable), the rest are summed; parameters no task touches get zeros.
"""
from __future__ import annotations

import torch

from .weight_methods import draw_visit_orders, pcgrad_merge


class PCGrad():
    def __init__(self, optimizer, reduction='mean'):
        self._optim, self._reduction = optimizer, reduction
        return

    @property
    def optimizer(self):
        return self._optim

    def zero_grad(self):
        return self._optim.zero_grad(set_to_none=True)

    def step(self):
        return self._optim.step()

    def _params(self):
        return [p for group in self._optim.param_groups for p in group['params']]

    def pc_backward(self, objectives):
        params = self._params()
        task_grads = []
        for obj in objectives:                                    # _pack_grad (:85-103)
            self._optim.zero_grad(set_to_none=True)
            obj.backward(retain_graph=True)
            task_grads.append([None if p.grad is None else p.grad.detach().clone() for p in params])
        T = len(task_grads)
        if self._reduction not in ('mean', 'sum'):
            exit('invalid reduction method')
        # a parameter nobody has a gradient for contributes zeros (the reference packs zeros_like, :127-130)
        zeros = {}
        for i, p in enumerate(params):
            if all(tg[i] is None for tg in task_grads):
                zeros[i] = torch.zeros_like(p)
                task_grads[0][i] = zeros[i]
        shared = [all(tg[i] is not None for tg in task_grads) and i not in zeros for i in range(len(params))]
        orders = draw_visit_orders(T)
        # coefficients are the SUM over tasks; shared entries are then averaged (scale 1/T)
        merged = pcgrad_merge(task_grads, orders, mean=False, seg_scales=[1.0 / T if s else 1.0 for s in shared])
        for p, g in zip(params, merged):                          # _set_grad (:72-83)
            p.grad = g
        return
