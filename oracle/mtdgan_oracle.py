"""CPU oracle for the MTD-GAN hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional torch-CPU / numpy code, the algorithm of the
reference hot path (babbu3682/MTD-GAN).  It is the checker the parity tests, the
`__graft_entry__.smoke()` check and the `cpu_baseline` / `--impl reference` legs of
`bench.py` use.  Nothing in the product path (`mtd-gan_b200/`, `arch/`, `losses.py`,
`module/`) may import it.

Parity status: PINNED.  `tests/test_oracle.py` checks every function below against
(a) the live reference imported from /root/reference (when present, i.e. in the build
container) and (b) the golden vectors committed under `tests/golden/`, which were
generated from the live reference by `tests/golden/make_golden.py`; the only
known-answer vector the reference itself ships (the `module/pcgrad.py:165-195` demo
printout) is checked in `tests/test_oracle.py::test_pcgrad_demo_known_answer`.

The arithmetic primitives (conv2d, rfft2, ...) are PyTorch's (third-party; the
reference pins torch==2.3.1 in requirements.txt:16, this image has 2.11.0), exactly as
in the reference; what is restated here is the reference's own composition of them.
All tensors are NCHW float32 like the reference's.

Weights are addressed by the reference's `state_dict` key names.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LEAK = 0.2   # nn.LeakyReLU(0.2), arch/Ours/networks.py:182 ff.


# --------------------------------------------------------------------------------------
# Res-FFT-Conv generator            (arch/Ours/networks.py:15-164)
# --------------------------------------------------------------------------------------
def fft_conv_block(x: Tensor, img_w: Tensor, img_b: Tensor, fft_w: Tensor, fft_b: Tensor) -> Tensor:
    """FFT_ConvBlock.forward, arch/Ours/networks.py:21-36."""
    H, W = x.shape[-2:]
    spec = torch.fft.rfft2(x, s=(H, W), dim=(2, 3), norm="ortho")          # :24
    stacked = torch.cat([spec.real, spec.imag], dim=1)                      # :25
    stacked = F.relu(F.conv2d(stacked, fft_w, fft_b))                       # :26
    re, im = torch.chunk(stacked, 2, dim=1)                                 # :27
    back = torch.fft.irfft2(torch.complex(re, im), s=(H, W), dim=(2, 3), norm="ortho")  # :28-29
    img = F.relu(F.conv2d(x, img_w, img_b, padding=1))                      # :32
    return x + img + back                                                   # :35


def _blk(sd: Dict[str, Tensor], i: int, x: Tensor, prefix: str = "") -> Tensor:
    p = f"{prefix}enforce.{i}."
    return fft_conv_block(x, sd[p + "img_conv.weight"], sd[p + "img_conv.bias"],
                          sd[p + "fft_conv.weight"], sd[p + "fft_conv.bias"])


def generator_forward(sd: Dict[str, Tensor], x: Tensor, prefix: str = "") -> Tensor:
    """ResFFT_Generator(1, 32, 10, 3, 1).forward, arch/Ours/networks.py:95-164.

    e_k = block_{k-1}(relu(enc_{k-1}(e_{k-1}))), k=1..10; bottleneck; then
    d_k = relu(dec_k(block(d_{k+1})) + e_k) with ConvTranspose2d(k=3,s=1,p=1) decoders.
    """
    def enc(i, t):
        return F.relu(F.conv2d(t, sd[f"{prefix}encoder.{i}.weight"], sd[f"{prefix}encoder.{i}.bias"], padding=1))

    def dec(i, t):
        return F.conv_transpose2d(t, sd[f"{prefix}decoder.{i}.weight"], sd[f"{prefix}decoder.{i}.bias"], padding=1)

    skips = []
    t = x
    for k in range(10):                       # :97-125
        t = _blk(sd, k, enc(k, t), prefix)
        skips.append(t)
    t = _blk(sd, 10, enc(10, t), prefix)      # :128-129
    t = F.relu(dec(10, t) + skips[9])         # :132
    for k in range(9, 0, -1):                 # :134-159   decoder[-(11-k)] == decoder[k]
        t = _blk(sd, 20 - k, t, prefix)
        t = F.relu(dec(k, t) + skips[k - 1])
    t = _blk(sd, 20, t, prefix)               # :161
    return F.relu(dec(0, t) + x)              # :162


# --------------------------------------------------------------------------------------
# Spectral norm (torch.nn.utils.spectral_norm semantics; call sites networks.py:181-300)
# --------------------------------------------------------------------------------------
def sn_power_iteration(w: Tensor, u: Tensor, v: Tensor, training: bool, eps: float = 1e-12
                       ) -> Tuple[Tensor, Tensor, Tensor]:
    """One power iteration (training) and sigma.  Returns (sigma, u_used, v_used).

    In training mode `u` and `v` are updated IN PLACE (like the reference's buffers) and
    clones are returned for the autograd graph.  SURVEY appendix A4.
    """
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
        u_, v_ = u.clone(), v.clone()
    else:
        u_, v_ = u, v
    sigma = torch.dot(u_, torch.mv(wm, v_))
    return sigma, u_, v_


# name, kind('c'=conv,'l'=linear), cin, cout, k, stride, pad, spectral_norm
def discriminator_layers(c: int = 64, cin: int = 1):
    """Layer table of Multi_Task_Discriminator_Skip(in_channels, out_channels=c) in
    registration order, arch/Ours/networks.py:181-306."""
    L = []
    chans = [(cin, c), (c, 2 * c), (2 * c, 4 * c), (4 * c, 8 * c), (8 * c, 8 * c), (8 * c, 8 * c)]
    for i, (a, b) in enumerate(chans, 1):
        L += [(f"conv{i}1", "c", a, b, 3, 1, 1, True), (f"conv{i}2", "c", b, b, 3, 1, 1, True),
              (f"down{i}", "c", b, b, 4, 2, 1, True)]
    L += [("bconv1", "c", 8 * c, 8 * c, 1, 1, 0, True), ("bconv2", "c", 8 * c, 8 * c, 1, 1, 0, True)]
    L += [("c_fc", "l", 512, 512, 1, 1, 0, True)]
    dec = [(16 * c, 8 * c), (16 * c, 8 * c), (16 * c, 4 * c), (8 * c, 2 * c), (4 * c, c), (2 * c, 1)]
    for i, (a, b) in enumerate(dec, 1):
        L += [(f"s_dconv{i}1", "c", a, b, 3, 1, 1, True), (f"s_dconv{i}2", "c", b, b, 3, 1, 1, True)]
    ups = [8 * c, 8 * c, 8 * c, 4 * c, 2 * c, c]
    for i, ((a, b), uc) in enumerate(zip(dec, ups), 1):
        L += [(f"r_up{i}.upsample.0", "c", uc, 4 * uc, 1, 1, 0, False),
              (f"r_dconv{i}1", "c", a, b, 3, 1, 1, True), (f"r_dconv{i}2", "c", b, b, 3, 1, 1, True)]
    L += [("enc_out", "l", 512, 1, 1, 1, 0, False), ("dec_out", "c", cin, 1, 1, 1, 0, False),
          ("rec_out", "c", cin, 1, 1, 1, 0, False)]
    return L


def discriminator_forward(sd: Dict[str, Tensor], x: Tensor, training: bool,
                          dropout_mask: Optional[Tensor] = None, prefix: str = ""
                          ) -> Tuple[Tensor, Tensor, Tensor]:
    """Multi_Task_Discriminator_Skip.forward, arch/Ours/networks.py:383-474.

    `sd` holds bias / weight_orig / weight_u / weight_v (spectral-normed layers) or
    weight / bias (plain layers).  In training mode every spectral-normed layer does one
    power iteration (u, v updated in place) before use.  `dropout_mask` is the
    already-scaled keep mask of shape (B, 512) (== F.dropout(ones, 0.3)); None means
    dropout is the identity (eval mode, or p effectively 0 for deterministic tests).
    """
    table = {n: (kind, k, s, p, sn) for n, kind, _, _, k, s, p, sn in discriminator_layers()}

    def weight(name):
        kind, k, s, p, sn = table[name]
        if not sn:
            return sd[f"{prefix}{name}.weight"]
        w = sd[f"{prefix}{name}.weight_orig"]
        sigma, _, _ = sn_power_iteration(w, sd[f"{prefix}{name}.weight_u"], sd[f"{prefix}{name}.weight_v"], training)
        return w / sigma

    def conv(name, t, act=True):
        kind, k, s, p, sn = table[name]
        t = F.conv2d(t, weight(name), sd[f"{prefix}{name}.bias"], stride=s, padding=p)
        return F.leaky_relu(t, LEAK) if act else t

    skips = []
    t = x
    for i in range(1, 7):                                    # :385-407 (no activation after down*)
        t = conv(f"conv{i}1", t)
        t = conv(f"conv{i}2", t)
        skips.append(t)
        t = conv(f"down{i}", t, act=False)
    t = conv("bconv1", t)                                    # :410
    x_bot = conv("bconv2", t)                                # :411

    h = x_bot.flatten(1)                                     # :414
    h = F.leaky_relu(F.linear(h, weight("c_fc"), sd[f"{prefix}c_fc.bias"]), LEAK)   # :415-416
    if dropout_mask is not None:                             # :417
        h = h * dropout_mask
    x_enc = F.linear(h, sd[f"{prefix}enc_out.weight"], sd[f"{prefix}enc_out.bias"])  # :470

    t = x_bot                                                # SEG decoder :420-442
    for i in range(1, 7):
        t = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
        t = conv(f"s_dconv{i}1", torch.cat([t, skips[6 - i]], dim=1))
        t = conv(f"s_dconv{i}2", t)
    x_dec = F.conv2d(t, sd[f"{prefix}dec_out.weight"], sd[f"{prefix}dec_out.bias"])  # :471

    t = x_bot                                                # REC decoder :445-467
    for i in range(1, 7):
        t = F.pixel_shuffle(F.conv2d(t, sd[f"{prefix}r_up{i}.upsample.0.weight"],
                                     sd[f"{prefix}r_up{i}.upsample.0.bias"]), 2)       # UpsampleBlock :166-175
        t = conv(f"r_dconv{i}1", torch.cat([t, skips[6 - i]], dim=1))
        t = conv(f"r_dconv{i}2", t)
    x_rec = F.conv2d(t, sd[f"{prefix}rec_out.weight"], sd[f"{prefix}rec_out.bias"])  # :472
    return x_enc, x_dec, x_rec


# --------------------------------------------------------------------------------------
# Losses                                  (losses.py:10-15, 99-138)
# --------------------------------------------------------------------------------------
def ls_gan(inputs: Tensor, targets: float) -> Tensor:
    """losses.py:10-11"""
    return torch.mean((inputs - targets) ** 2)


def nds_mask(diffs: Tensor) -> Tensor:
    """The boolean non-difference mask of NDS_Loss, losses.py:15 (`torch.abs(diffs).bool()`).
    0.0 / -0.0 -> False; denormals and NaN -> True."""
    return torch.abs(diffs).bool()


def nds_mask_numpy(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Bit-level restatement: (x - y) in IEEE fp32, non-zero test on the magnitude bits."""
    d = (x.astype(np.float32) - y.astype(np.float32)).view(np.uint32)
    return (d & np.uint32(0x7FFFFFFF)) != 0


def nds_loss(inputs: Tensor, targets: float, diffs: Tensor) -> Tensor:
    """losses.py:13-15 — mean over ALL elements."""
    return torch.mean(nds_mask(diffs) * (inputs - targets) ** 2)


def charbonnier(x: Tensor, y: Tensor, eps: float = 1e-3) -> Tensor:
    """losses.py:108-111"""
    d = x - y
    return torch.mean(torch.sqrt(d * d + eps * eps))


_GAUSS_1D = [0.05, 0.25, 0.4, 0.25, 0.05]     # losses.py:116


def _gauss_kernel(like: Optional[Tensor] = None) -> Tensor:
    k = torch.tensor([_GAUSS_1D], dtype=torch.float32)
    k = torch.matmul(k.t(), k)[None, None]      # losses.py:117 (1 gray channel)
    return k if like is None else k.to(device=like.device, dtype=like.dtype)   # losses.py:118-119 (kernel follows the device)


def _conv_gauss(img: Tensor) -> Tensor:
    """losses.py:122-125"""
    return F.conv2d(F.pad(img, (2, 2, 2, 2), mode="replicate"), _gauss_kernel(img))


def laplacian(img: Tensor) -> Tensor:
    """losses.py:127-134"""
    f = _conv_gauss(img)
    up = torch.zeros_like(f)
    up[:, :, ::2, ::2] = f[:, :, ::2, ::2] * 4
    return img - _conv_gauss(up)


def edge_loss(x: Tensor, y: Tensor) -> Tensor:
    """EdgeLoss.forward, losses.py:136-138"""
    return charbonnier(laplacian(x), laplacian(y))


# --------------------------------------------------------------------------------------
# Method wrapper                          (arch/Ours/networks.py:1957-2009)
# --------------------------------------------------------------------------------------
def d_loss(sd: Dict[str, Tensor], x: Tensor, y: Tensor, training: bool = True,
           dropout_masks: Optional[Sequence[Optional[Tensor]]] = None):
    """MTD_GAN_Method.d_loss — returns (Tensor[3], details dict of 10 scalars)."""
    dm = list(dropout_masks) if dropout_masks is not None else [None] * 4
    with torch.no_grad():
        fake = generator_forward(sd, x, "Generator.")                                      # :1958
    D = lambda t, m: discriminator_forward(sd, t, training, m, "Discriminator.")
    real_enc, real_dec, real_rec = D(y, dm[0])                                             # :1959
    fake_enc, fake_dec, fake_rec = D(fake, dm[1])                                          # :1960
    diff = x - y
    det = {"D/real_enc": ls_gan(real_enc, 1.0), "D/fake_enc": ls_gan(fake_enc, 0.0),
           "D/real_dec": nds_loss(real_dec, 1.0, diff), "D/fake_dec": nds_loss(fake_dec, 0.0, diff)}
    disc = det["D/real_enc"] + det["D/fake_enc"] + det["D/real_dec"] + det["D/fake_dec"]   # :1962
    det["D/rec_loss_real"] = F.l1_loss(real_rec, y)                                        # :1964
    det["D/rec_loss_fake"] = F.l1_loss(fake_rec, fake)                                     # :1965
    rec = det["D/rec_loss_real"] + det["D/rec_loss_fake"]
    rr_enc, rr_dec, _ = D(real_rec.clip(0, 1), dm[2])                                      # :1969
    rf_enc, rf_dec, _ = D(fake_rec.clip(0, 1), dm[3])                                      # :1970
    det["D/consist_loss_real_enc"] = F.mse_loss(real_enc, rr_enc)                          # :1972-1975
    det["D/consist_loss_real_dec"] = F.mse_loss(real_dec, rr_dec)
    det["D/consist_loss_fake_enc"] = F.mse_loss(fake_enc, rf_enc)
    det["D/consist_loss_fake_dec"] = F.mse_loss(fake_dec, rf_dec)
    consist = (det["D/consist_loss_real_enc"] + det["D/consist_loss_real_dec"]
               + det["D/consist_loss_fake_enc"] + det["D/consist_loss_fake_dec"])          # :1977
    return torch.stack([disc, rec, consist]), det                                          # :1992


def g_loss(sd: Dict[str, Tensor], x: Tensor, y: Tensor, training: bool = True,
           dropout_mask: Optional[Tensor] = None):
    """MTD_GAN_Method.g_loss — returns (scalar, details dict of 4 scalars)."""
    fake = generator_forward(sd, x, "Generator.")                                          # :1995
    gen_enc, gen_dec, _ = discriminator_forward(sd, fake, training, dropout_mask, "Discriminator.")
    diff = x - y
    det = {"G/gen_enc": ls_gan(gen_enc, 1.0), "G/gen_dec": nds_loss(gen_dec, 1.0, diff)}
    adv = det["G/gen_enc"] + det["G/gen_dec"]                                              # :1998
    det["G/pix_loss"] = 50.0 * charbonnier(fake, y)                                        # :1999
    det["G/edge_loss"] = 50.0 * edge_loss(fake, y)                                         # :2000
    return adv + det["G/pix_loss"] + det["G/edge_loss"], det                               # :2002


# --------------------------------------------------------------------------------------
# Parameter partitions                    (arch/Ours/networks.py:318-380)
# --------------------------------------------------------------------------------------
def d_shared_names() -> List[str]:
    """Parameter names (bias before weight_orig) of shared_parameters()."""
    out = []
    for i in range(1, 7):
        for n in (f"conv{i}1", f"conv{i}2", f"down{i}"):
            out += [f"{n}.bias", f"{n}.weight_orig"]
    for n in ("bconv1", "bconv2"):
        out += [f"{n}.bias", f"{n}.weight_orig"]
    return out


def d_task_specific_names() -> List[str]:
    out = []
    for i in range(1, 7):
        for n in (f"s_dconv{i}1", f"s_dconv{i}2"):
            out += [f"{n}.bias", f"{n}.weight_orig"]
    for i in range(1, 7):
        out += [f"r_up{i}.upsample.0.weight", f"r_up{i}.upsample.0.bias"]
        for n in (f"r_dconv{i}1", f"r_dconv{i}2"):
            out += [f"{n}.bias", f"{n}.weight_orig"]
    for n in ("enc_out", "dec_out", "rec_out"):
        out += [f"{n}.weight", f"{n}.bias"]
    return out


# --------------------------------------------------------------------------------------
# PCGrad                                  (module/weight_methods.py:449-464, module/pcgrad.py:50-70)
# --------------------------------------------------------------------------------------
def pcgrad_project_lists(grads: List[Tuple[Tensor, ...]], reduction: str = "sum",
                         rng: random.Random | None = None) -> List[Tensor]:
    """weight_methods.PCGrad._project_conflicting (:449-464) on per-parameter tensor lists.

    `grads` is shuffled IN PLACE once per outer task with Python's RNG (global `random`
    unless `rng` is given) — SURVEY Q6.  Dots are taken between the progressively projected
    g_i and the ORIGINAL g_j (including j == i).
    """
    shuffle = (rng or random).shuffle
    pc = [[g.clone() for g in task] for task in grads]
    for g_i in pc:
        shuffle(grads)
        for g_j in grads:
            dot = sum(torch.dot(a.flatten(), b.flatten()) for a, b in zip(g_i, g_j))
            if dot < 0:
                nsq = torch.norm(torch.cat([g.flatten() for g in g_j])) ** 2
                for a, b in zip(g_i, g_j):
                    a -= dot * b / nsq
    merged = [sum(g) for g in zip(*pc)]
    if reduction == "mean":
        merged = [g / len(pc) for g in merged]
    return merged


def pcgrad_project_flat(grads: List[Tensor], has_grads: List[Tensor],
                        rng: random.Random | None = None) -> Tensor:
    """module/pcgrad.py PCGrad._project_conflicting (:50-70) on flattened gradients.

    Shared entries (every task has a gradient) are ALWAYS averaged — the 'sum' branch at
    :63-65 is unreachable because `if self._reduction:` (:60) is truthy for both strings
    (SURVEY §3.3); non-shared entries are summed.
    """
    shuffle = (rng or random).shuffle
    shared = torch.stack(has_grads).prod(0).bool()
    pc = [g.clone() for g in grads]
    for g_i in pc:
        shuffle(grads)
        for g_j in grads:
            dot = torch.dot(g_i, g_j)
            if dot < 0:
                g_i -= dot * g_j / (g_j.norm() ** 2)
    merged = torch.zeros_like(grads[0])
    merged[shared] = torch.stack([g[shared] for g in pc]).mean(dim=0)
    merged[~shared] = torch.stack([g[~shared] for g in pc]).sum(dim=0)
    return merged


def pcgrad_coefficients(gram: np.ndarray, orders: Sequence[Sequence[int]]) -> np.ndarray:
    """Gram-space restatement: every projected g_i' stays in span{g_k}; returns C with
    g_i' = sum_k C[i,k] g_k, given the T x T Gram matrix and, for each outer task i, the
    order in which the ORIGINAL gradients g_j are visited (orders[i] = list of j)."""
    T = gram.shape[0]
    C = np.eye(T, dtype=np.float64)
    for i in range(T):
        for j in orders[i]:
            dot = float(C[i] @ gram[:, j])
            if dot < 0:
                C[i, j] -= dot / float(gram[j, j])
    return C


# --------------------------------------------------------------------------------------
# numpy statements of the FFT numerical contracts
# --------------------------------------------------------------------------------------
def irfft2_contract_numpy(re: np.ndarray, im: np.ndarray, H: int, W: int) -> np.ndarray:
    """irfft2(ortho) == complex iFFT along H, drop Im of columns kw=0 and kw=W/2, then the
    half-spectrum inverse along W (A1).  Independent of torch.fft."""
    spec = re.astype(np.float64) + 1j * im.astype(np.float64)
    col = np.fft.ifft(spec, axis=-2, norm="ortho")
    col[..., 0] = col[..., 0].real
    col[..., W // 2] = col[..., W // 2].real
    n = np.arange(W)
    k = np.arange(W // 2 + 1)
    wk = np.where((k == 0) | (k == W // 2), 1.0, 2.0)
    basis = np.exp(2j * np.pi * np.outer(k, n) / W)                 # (Wh, W)
    out = np.real((col * wk)[..., :, :, None] * basis).sum(axis=-2) / math.sqrt(W)
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# Synthetic benchmark inputs             
# --------------------------------------------------------------------------------------
def synthetic_pair(batch: int, size: int, seed: int = 1234, rank: int = 0) -> Tuple[Tensor, Tensor]:
    """(x, y): y = clamp(1.6*rand - 0.3, 0, 1) (plateaus at exactly 0 and 1, like HU windowing),
    x = clamp(y + 0.1*randn, 0, 1)  =>  ~19 % of pixels have x == y exactly (non-trivial NDS mask)."""
    g = torch.Generator().manual_seed(seed + rank)
    base = torch.rand(batch, 1, size, size, generator=g)
    y = torch.clamp(1.6 * base - 0.3, 0, 1)
    x = torch.clamp(y + 0.1 * torch.randn(batch, 1, size, size, generator=g), 0, 1)
    return x, y
