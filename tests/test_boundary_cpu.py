"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the modules reproduce the reference's state_dict / init / parameter partitions, CPU tensors
are rejected (no fallback), and the PCGrad host logic matches the oracle."""
import ctypes
import os
import random

import numpy as np
import pytest
import torch

from _golden_util import check_summary, load
from oracle import mtdgan_oracle as O


def test_library_exports_every_declared_symbol():
    import mtdgan_b200
    from mtdgan_b200 import _ext
    assert mtdgan_b200.is_built(), "libmtdgan_sm100a.so missing: run __graft_entry__.build()"
    protos = _ext.parse_header()
    assert len(protos) >= 35
    lib = ctypes.CDLL(_ext.lib_path())
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/mtdgan_b200.h but not exported"
    assert _ext.load().mtd_abi_version() == 1


def test_sass_is_sm100a_only():
    from mtdgan_b200 import _ext
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", _ext.lib_path()], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, archs


@pytest.fixture(scope="module")
def model():
    from arch.Ours.networks import MTD_GAN_Method
    torch.manual_seed(2024)
    random.seed(2024)
    return MTD_GAN_Method()


def test_init_fingerprint(model):
    fix = load("state_summary.pt")
    sd = model.state_dict()
    assert list(sd.keys()) == list(fix.keys())            # same 326 entries in the same order
    assert len(sd) == 326
    for k, s in fix.items():
        check_summary(sd[k], s, 1e-7, k)


def test_parameter_partitions(model):
    D, G = model.Discriminator, model.Generator
    names = {id(p): n for n, p in D.named_parameters()}
    shared = [names[id(p)] for p in D.shared_parameters()]
    ts = [names[id(p)] for p in D.task_specific_parameters()]
    assert shared == O.d_shared_names() and ts == O.d_task_specific_names()
    assert sum(p.numel() for p in D.shared_parameters()) == 28609920
    assert sum(p.numel() for p in D.task_specific_parameters()) == 39559451
    assert [names[id(p)] for p in D.last_shared_parameters()] == ["bconv2.bias", "bconv2.weight_orig"]
    assert not ({"c_fc.bias", "c_fc.weight_orig"} & set(shared + ts))      # Q1
    assert sum(p.numel() for p in G.parameters()) == 467137
    assert sum(p.numel() for p in D.parameters()) == 68432027
    assert len(list(G.shared_parameters())) == 44 and G.task_specific_parameters() is None   # Q7
    gl = {id(p): n for n, p in G.named_parameters()}
    assert [gl[id(p)] for p in G.last_shared_parameters()] == ["decoder.0.weight", "decoder.0.bias"]


def test_state_dict_roundtrip(model):
    from arch.Ours.networks import MTD_GAN_Method
    other = MTD_GAN_Method()
    other.load_state_dict(model.state_dict())
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k


def test_cpu_tensors_are_rejected(model):
    import losses
    from module.weight_methods import WeightMethods
    x = torch.zeros(1, 1, 64, 64)
    for fn in (lambda: model.Generator(x), lambda: model.Discriminator(x), lambda: model.d_loss(x, x),
               lambda: model.g_loss(x, x), lambda: losses.ls_gan(x, 1.0), lambda: losses.NDS_Loss(x, 1.0, x),
               lambda: losses.CharbonnierLoss()(x, x), lambda: losses.EdgeLoss()(x, x)):
        with pytest.raises(RuntimeError):
            fn()
    with pytest.raises(AssertionError):
        WeightMethods("nashmtl", n_tasks=3, device=torch.device("cpu"))


def test_visit_orders_consume_rng_like_reference():
    from mtdgan_b200.weight_methods import draw_visit_orders
    random.seed(123)
    orders = draw_visit_orders(3)
    after = random.random()
    random.seed(123)
    lst = ["a", "b", "c"]
    want = []
    for _ in range(3):
        random.shuffle(lst)
        want.append(["abc".index(s) for s in lst])
    assert orders == want and after == random.random()


def test_gram_space_pcgrad_matches_vector_space():
    rng = np.random.default_rng(0)
    for trial in range(50):
        T = 2 + trial % 3
        gs = [torch.tensor(rng.standard_normal(40), dtype=torch.float64) for _ in range(T)]
        if trial % 2:
            gs[1] = -gs[0] * 0.7 + 0.05 * gs[1]
        gram = np.array([[float(a @ b) for b in gs] for a in gs])
        r1, r2 = random.Random(trial), random.Random(trial)
        idx, orders = list(range(T)), []
        for _ in range(T):
            r1.shuffle(idx)
            orders.append(list(idx))
        C = O.pcgrad_coefficients(gram, orders)
        merged = sum(float(C.sum(0)[k]) * gs[k] for k in range(T))
        want = O.pcgrad_project_lists([(g.clone(),) for g in gs], "sum", rng=r2)[0]
        assert float((merged - want).norm() / want.norm()) <= 1e-10


def test_get_loss_and_module_surface(model):
    import losses
    assert isinstance(losses.get_loss("L1 Loss"), torch.nn.L1Loss)
    assert isinstance(losses.get_loss("L2 Loss"), torch.nn.MSELoss)
    with pytest.raises(Exception):
        losses.get_loss("nope")
    assert model.gan_metric_cls is losses.ls_gan and model.gan_metric_seg is losses.NDS_Loss
    assert tuple(model.edge_loss.kernel.shape) == (1, 1, 5, 5) and model.pixel_loss.eps == 1e-3
    for attr in ("conv11", "relu11", "down6", "bconv2", "brelu2", "c_flatten", "c_fc", "c_relu", "c_drop", "s_up1",
                 "s_dconv62", "s_drelu62", "r_up6", "r_dconv62", "r_drelu62", "enc_out", "dec_out", "rec_out"):
        assert hasattr(model.Discriminator, attr), attr


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = oracle port on the host cores): exactly one stdout line,
    a JSON object with the contract's keys (metric / unit / config shared with the GPU arm; impl, cpu_baseline, e2e)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("train patches/s") and d["unit"] == "patches/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_importing_bench_leaves_stdout_alone():
    """bench.py points fd 1 at stderr only in main(); its cpu_baseline leg imports the module in a child process and
    reads that child's stdout."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", "import bench; print('still-stdout')"], capture_output=True, text=True,
                       timeout=300, cwd=root)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "still-stdout", (r.stdout, r.stderr[-500:])
