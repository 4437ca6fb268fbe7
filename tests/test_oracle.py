"""CPU tests: the oracle restatement against vectors produced by the reference's own code (tests/golden)."""
import os

import pytest
import torch

from oracle import pea_oracle as O
from tests.util import load_golden, rel_err, state_dict_of

torch.set_num_threads(4)


def _adapter_case(name, fn, two_out, **kw):
    g = load_golden(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in state_dict_of(g).items()}
    x = g["x"].clone().requires_grad_(True)
    out = fn(sd, x, **kw)
    if two_out:
        x1, x2 = out
        assert rel_err(x1, g["x1"]) < 1e-6 and rel_err(x2, g["x2"]) < 1e-6
        ((x1 * g["g1"]).sum() + (x2 * g["g2"]).sum()).backward()
    else:
        assert rel_err(out, g["x2"]) < 1e-6
        (out * g["g2"]).sum().backward()
    assert rel_err(x.grad, g["dx"]) < 1e-5
    for k, v in sd.items():
        assert rel_err(v.grad, g["grad." + k]) < 1e-5, k


def test_adapter_sdxl_small():
    _adapter_case("adapter_sdxl_small", O.adapter_sdxl, True, use_residual=False)


def test_adapter_sdxl_residual():
    _adapter_case("adapter_sdxl_residual", O.adapter_sdxl, True, use_residual=True)


def test_adapter_sd_small():
    _adapter_case("adapter_sd_small", O.adapter_sd, False)


def test_adapter_sd_5linear():
    _adapter_case("adapter_sd5_small", O.adapter_sd, False)


def test_adapter_legacy():
    _adapter_case("adapter_legacy_small", O.adapter_legacy, True)


def test_adapter_sdxl_default_full_size():
    """The shipped 6.03 M-parameter adapter (train_sdxl_zh.py:107), weights regenerated from the seed."""
    g = load_golden("adapter_sdxl_default")
    sd = {k: v.requires_grad_(True) for k, v in O.adapter_init(1024, 1280, 1024, 2048, seed=0).items()}
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert sum(v.numel() for v in sd.values()) == 6033408
    x = torch.randn(2, 77, 1024, generator=torch.Generator().manual_seed(1234))
    x1, x2 = O.adapter_sdxl(sd, x)
    assert rel_err(x1, g["x1"]) < 1e-5
    assert rel_err(x2[:, ::7, ::16], g["x2_sub"]) < 1e-5
    assert abs(float(x2.detach().double().sum()) - float(g["x2_sum"])) < 1e-3 * max(1.0, abs(float(g["x2_sum"])))
    gg = torch.Generator().manual_seed(100)
    g1 = torch.randn(x1.shape, generator=gg)
    g2 = torch.randn(x2.shape, generator=gg)
    ((x1 * g1).sum() + (x2 * g2).sum()).backward()
    assert rel_err(sd["fc.bias"].grad, g["grad_fc_bias"]) < 1e-5
    assert rel_err(sd["layernorm.weight"].grad, g["grad_ln_weight"]) < 1e-4
    assert rel_err(sd["projector.0.weight"].grad[::32, ::32], g["grad_p0_sub"]) < 1e-4


def test_adapter_sd_default_full_size():
    g = load_golden("adapter_sd_default")
    sd = O.adapter_init(1024, 768, 2048, None, seed=0)
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert sum(v.numel() for v in sd.values()) == 7866368
    x = torch.randn(2, 77, 1024, generator=torch.Generator().manual_seed(1234))
    out = O.adapter_sd(sd, x)
    assert rel_err(out[:, ::7, ::8], g["x2_sub"]) < 1e-5


@pytest.mark.parametrize("name,skip", [("loss_sdxl", False), ("loss_sd_nanskip", True)])
def test_kd_loss_against_reference_statements(name, skip):
    g = load_golden(name)
    order = [str(k) for k in g["order"]]
    pred = g["pred"].clone().requires_grad_(True)
    fs = [g["s." + k].clone().requires_grad_(True) for k in order]
    ft = [g["t." + k] for k in order]
    total, l0, l1, l2 = O.kd_loss(pred, g["noise"], g["pred_t"], fs, ft, g["zh"], skip_nonfinite=skip)
    for a, b in ((total, g["total"]), (l0, g["l0"]), (l1, g["l1"]), (l2, g["l2"])):
        assert abs(float(a) - float(b)) <= 1e-6 * max(1.0, abs(float(b)))
    total.backward()
    assert rel_err(pred.grad, g["dpred"]) < 1e-6
    for k, t in zip(order, fs):
        ref = g["ds." + k]
        got = t.grad if t.grad is not None else torch.zeros_like(t)
        if float(ref.abs().sum()) == 0.0:
            assert float(got.abs().sum()) == 0.0, k  # the skipped (non-finite) tap
        else:
            assert rel_err(got, ref) < 1e-6, k


def test_kd_loss_hand_computed_toy():
    """2-sample toy with a value computed by hand."""
    pred = torch.tensor([1.0, 3.0]).view(2, 1, 1, 1)
    noise = torch.tensor([0.0, 1.0]).view(2, 1, 1, 1)
    pred_t = torch.tensor([2.0, 5.0]).view(2, 1, 1, 1)
    fs, ft = [torch.tensor([1.0, 1.0]).view(2, 1, 1, 1)], [torch.tensor([4.0, 9.0]).view(2, 1, 1, 1)]
    zh = torch.tensor([0, 1])
    total, l0, l1, l2 = O.kd_loss(pred, noise, pred_t, fs, ft, zh)
    assert float(l0) == pytest.approx((0 + 4.0) / 2)      # only sample 1 (zh=1): (3-1)^2
    assert float(l1) == pytest.approx((1.0 + 0) / 2)      # only sample 0: (1-2)^2
    assert float(l2) == pytest.approx((9.0 + 0) / 2)      # only sample 0: (4-1)^2
    assert float(total) == pytest.approx(2.0 + 0.5 + 0.45)


def test_attn2_sdpa_equals_explicit_fp64():
    g = torch.Generator().manual_seed(0)
    q = torch.randn(2, 50, 128, generator=g, dtype=torch.float64)
    k = torch.randn(2, 77, 128, generator=g, dtype=torch.float64)
    v = torch.randn(2, 77, 128, generator=g, dtype=torch.float64)
    assert rel_err(O.attn2_core(q, k, v, 2), O.attn2_core(q, k, v, 2, explicit=True)) < 1e-12
    assert rel_err(O.attn2_core(q.float(), k.float(), v.float(), 2), O.attn2_core(q, k, v, 2, explicit=True)) < 1e-5


def test_attn2_gradcheck_fp64():
    g = torch.Generator().manual_seed(1)
    mk = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64, requires_grad=True)
    h, ctx = mk(1, 5, 8), mk(1, 3, 6)
    wq, wk, wv, wo, bo = mk(8, 8), mk(8, 6), mk(8, 6), mk(8, 8), mk(8)
    assert torch.autograd.gradcheck(lambda *a: O.attn2(*a, heads=2, explicit=True), (h, ctx, wq, wk, wv, wo, bo))


def test_cfg_select_routes_rows():
    mask = torch.tensor([True, False, True])
    u, c = torch.zeros(3, 2, 4), torch.ones(3, 2, 4)
    out = O.cfg_select(mask, u, c)
    assert out[0].sum() == 0 and out[1].sum() == 8 and out[2].sum() == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference_class():
    """When the reference tree is present, run its MLP class live (AST-extracted) beside the oracle."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cls = mg.load_class("train_sdxl_zh.py", "MLP")
    torch.manual_seed(5)
    m = cls(32, 40, 48, 56, False)
    x = torch.randn(3, 9, 32)
    x1, x2 = m(x)
    o1, o2 = O.adapter_sdxl(dict(m.state_dict()), x)
    assert torch.allclose(x1, o1, atol=1e-6) and torch.allclose(x2, o2, atol=1e-6)
