"""The plain-C oracle (oracle/pea_oracle_c.c) against the reference-generated golden vectors and against the PyTorch
oracle: two restatements that share no code must agree, and both must reproduce what the reference's own classes and
loss statements produced (tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as OC
from oracle import pea_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _weights(g):
    idx = sorted(int(k.split(".")[2]) for k in g.files if k.startswith("sd.projector.") and k.endswith(".weight"))
    return [g[f"sd.projector.{i}.weight"] for i in idx]


@pytest.mark.parametrize("name", ["adapter_sdxl_small", "adapter_sdxl_residual"])
def test_c_adapter_sdxl_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    res = bool(g["use_residual"]) if "use_residual" in g.files else False
    x1, x2 = OC.adapter(g["x"], g["sd.layernorm.weight"], g["sd.layernorm.bias"], _weights(g), g["sd.fc.weight"], g["sd.fc.bias"], res)
    assert _rel(x1, g["x1"]) < 2e-6 and _rel(x2, g["x2"]) < 2e-6


@pytest.mark.parametrize("name", ["adapter_sd_small", "adapter_sd5_small"])
def test_c_adapter_sd_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    y = OC.adapter(g["x"], g["sd.layernorm.weight"], g["sd.layernorm.bias"], _weights(g))
    assert _rel(y, g["x2"]) < 2e-6


def test_c_adapter_shipped_size_matches_reference_golden():
    """The shipped 6.03 M-parameter SDXL adapter (train_sdxl_zh.py:107) on [2, 77, 1024]: weights and input regenerated from
    the seeds tests/golden/make_golden.py used; the golden holds x1, a strided sample of x2 and its sum."""
    g = np.load(os.path.join(GOLD, "adapter_sdxl_default.npz"))
    sd = O.adapter_init(1024, 1280, 1024, 2048, seed=0)
    x = torch.randn(2, 77, 1024, generator=torch.Generator().manual_seed(1234))
    ws = [sd[f"projector.{i}.weight"].numpy() for i in (0, 2, 4)]
    x1, x2 = OC.adapter(x.numpy(), sd["layernorm.weight"].numpy(), sd["layernorm.bias"].numpy(), ws, sd["fc.weight"].numpy(),
                        sd["fc.bias"].numpy())
    assert _rel(x1, g["x1"]) < 1e-5 and _rel(x2[:, ::7, ::16], g["x2_sub"]) < 1e-5
    assert abs(float(x2.astype(np.float64).sum()) - float(g["x2_sum"])) < 1e-3 * max(1.0, abs(float(g["x2_sum"])))


@pytest.mark.parametrize("name,skip", [("loss_sdxl", False), ("loss_sd_nanskip", True)])
def test_c_kd_loss_matches_reference_golden(name, skip):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    order = [str(k) for k in g["order"]]
    (tot, l0, l1, l2), dpred, dfs = OC.kd_loss(g["pred"], g["noise"], g["pred_t"], [g["s." + k] for k in order],
                                               [g["t." + k] for k in order], g["zh"], 0.1, skip)
    for got, key in ((tot, "total"), (l0, "l0"), (l1, "l1"), (l2, "l2")):
        assert abs(got - float(g[key])) <= 2e-6 * max(1.0, abs(float(g[key]))), key
    assert _rel(dpred, g["dpred"]) < 2e-6
    for k, d in zip(order, dfs):
        ref = g["ds." + k]
        if np.abs(ref).max() == 0:           # a skipped (non-finite) tap carries no gradient
            assert np.abs(d).max() == 0, k
        else:
            assert _rel(d, ref) < 2e-6, k


@pytest.mark.parametrize("B,N,L,H,d", [(2, 37, 77, 3, 64), (1, 5, 1, 2, 40), (2, 16, 52, 2, 160)])
def test_c_attn2_core_matches_python_oracle(B, N, L, H, d):
    gen = torch.Generator().manual_seed(B * 1000 + N)
    C = H * d
    q = (torch.randn(B, N, C, generator=gen) * 2).double().requires_grad_(True)
    k = torch.randn(B, L, C, generator=gen).double().requires_grad_(True)
    v = torch.randn(B, L, C, generator=gen).double().requires_grad_(True)
    do = torch.randn(B, N, C, generator=gen).double()
    o = O.attn2_core(q, k, v, H, explicit=True)
    o.backward(do)
    f = lambda t: t.detach().float().numpy()
    assert _rel(OC.attn2_core(f(q), f(k), f(v), H), o.detach().numpy()) < 1e-6
    dq, dk, dv = OC.attn2_core_bwd(f(q), f(k), f(v), f(do), H)
    assert _rel(dv, v.grad.numpy()) < 1e-6
    if L > 1:                                # one key: P == 1, dQ = dK = 0 exactly
        assert _rel(dq, q.grad.numpy()) < 1e-6 and _rel(dk, k.grad.numpy()) < 1e-6
    else:
        assert np.abs(dq).max() < 1e-6 and np.abs(dk).max() < 1e-6


def test_c_and_python_adapter_agree_on_seeded_ragged_case():
    torch.manual_seed(3)
    sd = O.adapter_init(40, 56, 72, 88)
    x = torch.randn(3, 11, 40)
    x1, x2 = O.adapter_sdxl(sd, x)
    ws = [sd[k].numpy() for k in sorted((k for k in sd if k.startswith("projector.")), key=lambda s: int(s.split(".")[1]))]
    c1, c2 = OC.adapter(x.numpy(), sd["layernorm.weight"].numpy(), sd["layernorm.bias"].numpy(), ws, sd["fc.weight"].numpy(),
                        sd["fc.bias"].numpy())
    assert _rel(c1, x1.numpy()) < 2e-6 and _rel(c2, x2.numpy()) < 2e-6
