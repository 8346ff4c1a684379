"""Loss terms of MTD-GAN on the B200 path.  Public names and call signatures mirror the reference's
losses.py (ls_gan :10-11, NDS_Loss :13-15, CharbonnierLoss :99-111, EdgeLoss :113-138, get_loss
:186-197); the grouped *_terms functions are what MTD_GAN_Method.d_loss / g_loss use (each group is a
single autograd node returning [total, term_1, ...]).

Every reduction is a CUDA kernel accumulating in fp64 on the device; nothing synchronises with the
host.  Inputs of any shape are treated as flat fp32 arrays (the reference's NCHW tensors have C == 1
here, so NCHW and NHWC coincide).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _ext
from ._ext import call, fptr, ptr, stream
from .ops import check_input

L1, SQ, CHARB = 0, 1, 2


def _acc(k, like):
    return torch.zeros(k, dtype=torch.float64, device=like.device)


def _finalize(acc, scales, like):
    k = len(scales)
    out = torch.empty(1 + k, dtype=torch.float32, device=like.device)
    s = list(scales) + [0.0] * (4 - k)
    call("mtd_loss_finalize", ptr(acc), k, s[0], s[1], s[2], s[3], fptr(out), stream())
    return out


def _c(t):
    return t.detach().contiguous()


class _SqErrTerms(Function):
    """out = [sum_k term_k, term_1..term_K], term_k = mean(mask_k * (in_k - target_k)^2).
    spec: tuple of (target, masked) per input; x, y give the NDS mask (losses.py:15)."""

    @staticmethod
    def forward(ctx, x, y, spec, *inputs):
        st = stream()
        acc = _acc(len(inputs), inputs[0])
        ins = [_c(t) for t in inputs]
        xs, ys = (_c(x), _c(y)) if x is not None else (None, None)
        scales = []
        for k, (t, (target, masked)) in enumerate(zip(ins, spec)):
            if masked:
                assert t.numel() == xs.numel()
            call("mtd_sum_sqerr", fptr(t), float(target), fptr(xs) if masked else None, fptr(ys) if masked else None,
                 t.numel(), acc.data_ptr() + 8 * k, st)
            scales.append(1.0 / t.numel())
        ctx.spec, ctx.scales = spec, scales
        ctx.save_for_backward(xs, ys, *ins)
        return _finalize(acc, scales, ins[0])

    @staticmethod
    def backward(ctx, gout):
        xs, ys, *ins = ctx.saved_tensors
        gout = gout.contiguous()
        st = stream()
        grads = []
        for k, (t, (target, masked)) in enumerate(zip(ins, ctx.spec)):
            if not ctx.needs_input_grad[3 + k]:
                grads.append(None)
                continue
            d = torch.empty_like(t)
            call("mtd_sqerr_bwd", fptr(t), float(target), fptr(xs) if masked else None, fptr(ys) if masked else None,
                 t.numel(), fptr(gout), 1 + k, ctx.scales[k], fptr(d), st)
            grads.append(d)
        return (None, None, None, *grads)


class _DiffTerms(Function):
    """out = [sum_k term_k, term_1..], term_k = weight_k * mean(f(a_k - b_k)), f = |.|, (.)^2 or
    Charbonnier.  inputs = a_1, b_1, a_2, b_2, ...; gradients flow to both members of each pair."""

    @staticmethod
    def forward(ctx, mode, eps, weights, *pairs):
        st = stream()
        n = len(pairs) // 2
        acc = _acc(n, pairs[0])
        ts = [_c(t) for t in pairs]
        scales = []
        for k in range(n):
            a, b = ts[2 * k], ts[2 * k + 1]
            assert a.numel() == b.numel()
            call("mtd_sum_diff", fptr(a), fptr(b), a.numel(), mode, float(eps), acc.data_ptr() + 8 * k, st)
            scales.append(weights[k] / a.numel())
        ctx.mode, ctx.eps, ctx.scales = mode, eps, scales
        ctx.save_for_backward(*ts)
        return _finalize(acc, scales, ts[0])

    @staticmethod
    def backward(ctx, gout):
        ts = ctx.saved_tensors
        gout = gout.contiguous()
        st = stream()
        grads = []
        for k in range(len(ts) // 2):
            a, b = ts[2 * k], ts[2 * k + 1]
            na, nb = ctx.needs_input_grad[3 + 2 * k], ctx.needs_input_grad[4 + 2 * k]
            da = torch.empty_like(a) if na else None
            db = torch.empty_like(b) if nb else None
            if na or nb:
                call("mtd_diff_bwd", fptr(a), fptr(b), a.numel(), ctx.mode, float(ctx.eps), fptr(gout), 1 + k, ctx.scales[k],
                     fptr(da), fptr(db), st)
            grads += [da, db]
        return (None, None, None, *grads)


class _EdgeTerm(Function):
    """[edge] = weight * mean sqrt(lap(x - y)^2 + eps^2) over (B,1,H,W) images (losses.py:136-138)."""

    @staticmethod
    def forward(ctx, x, y, eps, weight):
        xs, ys = _c(x), _c(y)
        B, H, W = xs.shape[0], xs.shape[-2], xs.shape[-1]
        if xs.numel() != B * H * W:
            raise _ext.MtdError("EdgeLoss expects single-channel images (B,1,H,W)")
        acc = _acc(1, xs)
        call("mtd_sum_edge", fptr(xs), fptr(ys), B, H, W, float(eps), ptr(acc), stream())
        ctx.dims, ctx.eps, ctx.scale = (B, H, W), eps, weight / xs.numel()
        ctx.save_for_backward(xs, ys)
        return _finalize(acc, [ctx.scale], xs)

    @staticmethod
    def backward(ctx, gout):
        xs, ys = ctx.saved_tensors
        B, H, W = ctx.dims
        gout = gout.contiguous()
        dx = torch.empty_like(xs)
        call("mtd_edge_bwd", fptr(xs), fptr(ys), B, H, W, float(ctx.eps), fptr(gout), 1, ctx.scale, -1, 0.0, fptr(dx), stream())
        dy = -dx if ctx.needs_input_grad[1] else None
        return dx, dy, None, None


class _GLossTerms(Function):
    """[total, gen_enc, gen_dec, pix, edge] of MTD_GAN_Method.g_loss (arch/Ours/networks.py:1998-2007)
    in one node; the Charbonnier and Edge gradients w.r.t. `fake` are produced by one kernel."""

    @staticmethod
    def forward(ctx, gen_enc, gen_dec, fake, x, y, eps, w_pix, w_edge):
        st = stream()
        ge, gd, fk, xs, ys = _c(gen_enc), _c(gen_dec), _c(fake), _c(x), _c(y)
        B, H, W = fk.shape[0], fk.shape[-2], fk.shape[-1]
        acc = _acc(4, fk)
        call("mtd_sum_sqerr", fptr(ge), 1.0, None, None, ge.numel(), acc.data_ptr(), st)
        call("mtd_sum_sqerr", fptr(gd), 1.0, fptr(xs), fptr(ys), gd.numel(), acc.data_ptr() + 8, st)
        call("mtd_sum_diff", fptr(fk), fptr(ys), fk.numel(), CHARB, float(eps), acc.data_ptr() + 16, st)
        call("mtd_sum_edge", fptr(fk), fptr(ys), B, H, W, float(eps), acc.data_ptr() + 24, st)
        scales = [1.0 / ge.numel(), 1.0 / gd.numel(), w_pix / fk.numel(), w_edge / fk.numel()]
        ctx.scales, ctx.eps, ctx.dims = scales, eps, (B, H, W)
        ctx.save_for_backward(ge, gd, fk, xs, ys)
        return _finalize(acc, scales, fk)

    @staticmethod
    def backward(ctx, gout):
        ge, gd, fk, xs, ys = ctx.saved_tensors
        B, H, W = ctx.dims
        gout = gout.contiguous()
        st = stream()
        need = ctx.needs_input_grad
        dge = dgd = dfk = None
        if need[0]:
            dge = torch.empty_like(ge)
            call("mtd_sqerr_bwd", fptr(ge), 1.0, None, None, ge.numel(), fptr(gout), 1, ctx.scales[0], fptr(dge), st)
        if need[1]:
            dgd = torch.empty_like(gd)
            call("mtd_sqerr_bwd", fptr(gd), 1.0, fptr(xs), fptr(ys), gd.numel(), fptr(gout), 2, ctx.scales[1], fptr(dgd), st)
        if need[2]:
            dfk = torch.empty_like(fk)
            call("mtd_edge_bwd", fptr(fk), fptr(ys), B, H, W, float(ctx.eps), fptr(gout), 4, ctx.scales[3], 3, ctx.scales[2],
                 fptr(dfk), st)
        return dge, dgd, dfk, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# grouped terms used by MTD_GAN_Method
# ------------------------------------------------------------------------------------------------
def disc_terms(real_enc, fake_enc, real_dec, fake_dec, x, y):
    """[disc, real_enc, fake_enc, real_dec, fake_dec]  (networks.py:1962, 1979-1982)"""
    return _SqErrTerms.apply(x, y, ((1.0, False), (0.0, False), (1.0, True), (0.0, True)), real_enc, fake_enc, real_dec,
                             fake_dec)


def rec_terms(real_rec, y, fake_rec, fake):
    """[rec, l1(real_rec,y), l1(fake_rec,fake)]  (networks.py:1964-1966)"""
    return _DiffTerms.apply(L1, 0.0, (1.0, 1.0), real_rec, y, fake_rec, fake)


def consist_terms(real_enc, rr_enc, real_dec, rr_dec, fake_enc, rf_enc, fake_dec, rf_dec):
    """[consist, real_enc, real_dec, fake_enc, fake_dec]  (networks.py:1972-1977)"""
    return _DiffTerms.apply(SQ, 0.0, (1.0, 1.0, 1.0, 1.0), real_enc, rr_enc, real_dec, rr_dec, fake_enc, rf_enc, fake_dec,
                            rf_dec)


def g_terms(gen_enc, gen_dec, fake, x, y, eps=1e-3, w_pix=50.0, w_edge=50.0):
    """[total, gen_enc, gen_dec, 50*pix, 50*edge]  (networks.py:1998-2007)"""
    return _GLossTerms.apply(gen_enc, gen_dec, fake, x, y, eps, w_pix, w_edge)


# ------------------------------------------------------------------------------------------------
# reference-compatible public surface (losses.py)
# ------------------------------------------------------------------------------------------------
def _target_value(targets) -> float:
    if isinstance(targets, torch.Tensor):
        if targets.numel() != 1:
            raise _ext.MtdError("ls_gan / NDS_Loss on the B200 path take a scalar target (the reference passes 0. or 1.)")
        return float(targets)
    return float(targets)


def ls_gan(inputs, targets):
    """losses.py:10-11 — torch.mean((inputs - targets) ** 2)"""
    check_input(inputs, "ls_gan")
    return _SqErrTerms.apply(None, None, ((_target_value(targets), False),), inputs)[0]


def NDS_Loss(inputs, targets, diffs):
    """losses.py:13-15 — mean(|diffs|.bool() * (inputs - targets)**2); `diffs` is x - y in the reference's
    only call sites; the mask kernel takes the pair (diffs, 0)."""
    check_input(inputs, "NDS_Loss")
    check_input(diffs, "NDS_Loss")
    zero = torch.zeros_like(diffs)
    return _SqErrTerms.apply(diffs, zero, ((_target_value(targets), True),), inputs)[0]


def nds_mask(x, y):
    """The boolean NDS mask |x - y| != 0, bit-exact with torch.abs(x - y).bool()."""
    xs, ys = check_input(x, "nds_mask"), check_input(y, "nds_mask")
    m = torch.empty(xs.shape, dtype=torch.uint8, device=xs.device)
    call("mtd_nds_mask", fptr(xs), fptr(ys), ptr(m), xs.numel(), stream())
    return m.bool()


class CharbonnierLoss(torch.nn.Module):
    """losses.py:99-111"""

    def __init__(self, eps=1e-3):
        super().__init__()
        self.eps = eps

    def forward(self, x, y):
        check_input(x, "CharbonnierLoss")
        check_input(y, "CharbonnierLoss")
        return _DiffTerms.apply(CHARB, self.eps, (1.0,), x, y)[0]


class EdgeLoss(torch.nn.Module):
    """losses.py:113-138.  `kernel` (the 5x5 Gaussian) is kept as an attribute like the reference's; the
    CUDA kernel carries the same taps as constants."""

    def __init__(self):
        super().__init__()
        k = torch.Tensor([[.05, .25, .4, .25, .05]])
        self.kernel = torch.matmul(k.t(), k).unsqueeze(0).repeat(1, 1, 1, 1)
        if torch.cuda.is_available():
            self.kernel = self.kernel.cuda()
        self.loss = CharbonnierLoss()

    def forward(self, x, y):
        check_input(x, "EdgeLoss")
        check_input(y, "EdgeLoss")
        return _EdgeTerm.apply(x, y, self.loss.eps, 1.0)[0]


def get_loss(name):
    """losses.py:186-197 (validation/test criterion; outside the hot path, plain torch modules)."""
    if name == 'L2 Loss':
        return torch.nn.MSELoss()
    if name == 'L1 Loss':
        return torch.nn.L1Loss()
    raise Exception('Error...! name')
