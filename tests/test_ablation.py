"""Ablation rows (reference arch/Ours/networks.py:478-1936, models.py:56-75): construction parity on CPU
(state_dict keys, shapes and same-seed initial values against fingerprints of the live reference) and, on the GPU,
`d_loss` / `g_loss` totals, details, gradients and spectral-norm buffers against tests/golden/ablation.pt
(generated from the live reference by tests/golden/make_golden_ablation.py)."""
import random

import pytest
import torch

from _golden_util import GradTally, check_summary, load, rel_err
from oracle import mtdgan_oracle as O

NAMES = ["Ablation_CLS", "Ablation_SEG", "Ablation_CLS_SEG", "Ablation_CLS_REC", "Ablation_SEG_REC", "Ablation_CLS_SEG_REC",
         "Ablation_CLS_SEG_REC_NDS", "Ablation_CLS_SEG_REC_RC", "Ablation_CLS_SEG_REC_NDS_RC",
         "Ablation_CLS_SEG_REC_NDS_RC_ResFFT"]


def build(name):
    import arch.Ours.networks as N
    torch.manual_seed(2024)
    random.seed(2024)
    return getattr(N, name)()


@pytest.mark.parametrize("name", NAMES)
def test_construction_matches_reference(name):
    """Same seed => same state_dict keys (order included), shapes and initial values as the reference class."""
    fix = load("ablation.pt")[name]
    m = build(name)
    sd = m.state_dict()
    assert list(sd.keys()) == list(fix["state"].keys())
    for k, v in sd.items():
        check_summary(v, fix["state"][k], 0.0, k)          # bit-identical init (same RNG consumption order)


def test_models_py_names_all_construct():
    """Every model name of the reference's `get_model` ablation branch (models.py:56-75) resolves in the drop-in."""
    import arch.Ours.networks as N
    for name in NAMES + ["MTD_GAN_Method"]:
        assert callable(getattr(N, name))
    assert callable(N.REDCNN_Generator) and callable(N.SEG_REC_Discriminator)


def test_cpu_tensors_rejected_by_ablation_modules():
    import arch.Ours.networks as N
    from mtdgan_b200._ext import MtdError
    with pytest.raises((MtdError, RuntimeError)):
        N.REDCNN_Generator(1, 32, 10, 3, 1)(torch.zeros(1, 1, 64, 64))


def drop_mask(b, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(b, 512, generator=g) >= 0.3).float() / 0.7


CASES = [(n, "tc3") for n in NAMES] + [("Ablation_CLS", "simt"), ("Ablation_CLS_SEG_REC", "simt"), ("Ablation_CLS_SEG_REC_NDS_RC", "simt")]


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", CASES, ids=[f"{n}-{m}" for n, m in CASES])
def test_ablation_losses_and_gradients_vs_golden(name, mode):
    """tc3: tcgen05 3xTF32 forward / dgrad + plain-TF32 weight gradients (the default); simt: exact fp32 kernels."""
    from mtdgan_b200 import networks as NW, ops
    ops.set_conv_mode("simt" if mode == "simt" else "auto", 3)
    try:
        _ablation_case(name, mode)
    finally:
        ops.set_conv_mode("auto", 3)


def _ablation_case(name, mode):
    from mtdgan_b200 import networks as NW
    fix = load("ablation.pt")[name]
    m = build(name).to("cuda").train()
    x, y = (t.to("cuda") for t in O.synthetic_pair(2, 64, seed=77))
    q = [drop_mask(2, 900 + i).to("cuda") for i in range(5)]
    NW.set_dropout_mask_provider(lambda b, n, dev: q.pop(0))
    # tc3: 3xTF32 forward / dgrad (3e-4 on outputs), plain-TF32 weight gradients (2e-3 class); simt: fp32 (1e-4 / 1e-3)
    tol_out, tol_grad = (3e-4, 2e-3) if mode == "tc3" else (1e-4, 1e-3)
    try:
        d_total, d_det = m.d_loss(x, y)
        assert d_total.dim() == 0
        d_total.backward()
        assert abs(float(d_total) - fix["d_total"]) <= tol_out * abs(fix["d_total"])
        assert list(d_det.keys()) == list(fix["d_details"].keys())
        for k, v in fix["d_details"].items():
            assert abs(float(d_det[k]) - v) <= (10 if v > 1e-3 else 300) * tol_out * abs(v) + 1e-10, k
        # tc3: the shared-encoder gradients of the 1- and 3-head discriminators at B = 2 are limited by LeakyReLU mask flips
        # under the ~1e-5 forward perturbation of 3xTF32 (100 x fp32's: the fp32-vs-fp64 floor of these tensors is
        # 2.5e-5, theirs 2e-3 .. 1e-2); the exact-fp32 mode of the same kernels holds 1e-3 on every tensor (CASES),
        # so in tc3 mode the bulk must meet 2e-3 and nothing may exceed the gross-error bound
        tally = GradTally(frac_outliers=0.40 if mode == "tc3" else 0.04, gross=5e-2)
        for k, p in m.Discriminator.named_parameters():
            if fix["d_grads"][k] is None:
                assert p.grad is None, k
            else:
                check_summary(p.grad, fix["d_grads"][k], tol_grad, k, noise=fix["d_grads_noise"][k], tally=tally)
        tally.finish(2e-3)
        m.zero_grad(set_to_none=True)
        g_total, g_det = m.g_loss(x, y)
        g_total.backward()
        assert abs(float(g_total) - fix["g_total"]) <= tol_out * abs(fix["g_total"])
        assert list(g_det.keys()) == list(fix["g_details"].keys())
        for k, v in fix["g_details"].items():
            assert abs(float(g_det[k]) - v) <= 10 * tol_out * abs(v) + 1e-8, k
        tally = GradTally()
        for k, p in m.Generator.named_parameters():
            check_summary(p.grad, fix["g_grads"][k], tol_grad, k, noise=fix["g_grads_noise"][k], tally=tally)
        tally.finish()
        for k, v in m.Discriminator.named_buffers():
            check_summary(v, fix["buffers"][k], 1e-4, k)
    finally:
        NW.set_dropout_mask_provider(None)


@pytest.mark.gpu
def test_redcnn_generator_forward_vs_golden():
    import arch.Ours.networks as N
    torch.manual_seed(2024)
    G = N.REDCNN_Generator(in_channels=1, out_channels=32, num_layers=10, kernel_size=3, padding=1).to("cuda").eval()
    with torch.no_grad():
        out = G(O.synthetic_pair(2, 64, seed=11)[0].to("cuda"))
    assert rel_err(out, load("ablation.pt")["redcnn_fwd_64"]) <= 1e-4
