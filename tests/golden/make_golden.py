"""Generate the committed golden vectors by executing the REFERENCE's own code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

* Adapter classes: the ``ClassDef`` nodes named ``MLP`` / ``MLP_plus`` are extracted with ``ast`` from the
  reference scripts and ``exec``-ed verbatim (they depend on torch only; the scripts themselves cannot be imported
  because pytorch_lightning / diffusers / cn_clip are absent -- SURVEY.md section 8c).
* KD loss: the loss statements of ``training_step`` are extracted BY LINE RANGE from the reference scripts, dedented,
  stripped of the Lightning ``self.log`` / optimizer-lr lines, and exec-ed against stub ``self`` / ``args`` / ``batch``
  objects.  No reference source text is written into this repository -- only the resulting tensors.

Outputs: tests/golden/*.npz (small).  Nothing under tests/ reads /root/reference at run time.
"""
from __future__ import annotations

import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PEA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pea_oracle as O  # noqa: E402  (only for adapter_init: seeded full-size weights)


def load_class(relpath: str, name: str):
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    ns = {"torch": torch, "nn": nn, "F": F}
    exec(compile(ast.Module(body=[node], type_ignores=[]), relpath, "exec"), ns)
    return ns[name]


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in out.items() if np.ndim(v) > 0 and np.size(v) > 64})


def run_adapter(cls, ctor_args, x, seed, two_out, sd_override=None):
    torch.manual_seed(seed)
    m = cls(*ctor_args)
    if sd_override is not None:
        m.load_state_dict(sd_override, strict=True)
    # make LayerNorm affine non-trivial so gamma/beta are exercised
    elif True:
        with torch.no_grad():
            m.layernorm.weight.uniform_(0.5, 1.5)
            m.layernorm.bias.uniform_(-0.2, 0.2)
    x = x.clone().requires_grad_(True)
    out = m(x)
    g = torch.Generator().manual_seed(seed + 100)
    if two_out:
        x1, x2 = out
        g1 = torch.randn(x1.shape, generator=g)
        g2 = torch.randn(x2.shape, generator=g)
        ((x1 * g1).sum() + (x2 * g2).sum()).backward()
        outs = {"x1": x1, "x2": x2, "g1": g1, "g2": g2}
    else:
        g2 = torch.randn(out.shape, generator=g)
        (out * g2).sum().backward()
        outs = {"x2": out, "g2": g2}
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    grads = {"grad." + k: p.grad.detach().clone() for k, p in m.named_parameters()}
    return m, sd, outs, grads, x.grad.detach().clone()


def adapters():
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 7, 64, generator=gen)
    # A1 SDXL (train_sdxl_zh.py:43-67)
    cls = load_class("train_sdxl_zh.py", "MLP")
    for tag, args in (("sdxl_small", (64, 80, 96, 128, False)), ("sdxl_residual", (64, 64, 96, 128, True))):
        m, sd, outs, grads, dx = run_adapter(cls, args, x, 0, True)
        npz("adapter_" + tag, ctor=np.array(args[:4]), use_residual=args[4], x=x, dx=dx,
            **{"sd." + k: v for k, v in sd.items()}, **outs, **grads)
    # same class as shipped in the inference scripts must agree with the training one
    cls_t = load_class("tests/test_sdxl_zh.py", "MLP")
    cls_p = load_class("tests/test_sdxl_zh_lcm.py", "MLP_plus")
    torch.manual_seed(0); a = cls(64, 80, 96, 128, False)
    for other in (cls_t, cls_p):
        b = other(64, 80, 96, 128, False); b.load_state_dict(a.state_dict(), strict=True)
        assert all(torch.equal(p, q) for p, q in zip(a(x), b(x)))
    # A2 SD1.5 (train_sd_zh.py:41-56)
    cls = load_class("train_sd_zh.py", "MLP")
    m, sd, outs, grads, dx = run_adapter(cls, (64, 48, 96), x, 1, False)
    npz("adapter_sd_small", ctor=np.array((64, 48, 96)), x=x, dx=dx, **{"sd." + k: v for k, v in sd.items()}, **outs, **grads)
    # A4 SD test variant, 5 Linears (tests/test_sd_zh.py:25-44)
    cls = load_class("tests/test_sd_zh.py", "MLP")
    m, sd, outs, grads, dx = run_adapter(cls, (64, 48, 96), x, 2, False)
    npz("adapter_sd5_small", ctor=np.array((64, 48, 96)), x=x, dx=dx, **{"sd." + k: v for k, v in sd.items()}, **outs, **grads)
    # A4 legacy biased (tests/test_sdxl_zh_lcm.py:64-87)
    cls = load_class("tests/test_sdxl_zh_lcm.py", "MLP")
    m, sd, outs, grads, dx = run_adapter(cls, (64, 80, 96, 128, False), x, 3, True)
    npz("adapter_legacy_small", ctor=np.array((64, 80, 96, 128)), x=x, dx=dx, **{"sd." + k: v for k, v in sd.items()}, **outs, **grads)

    # shipped default A1 (1024,1280,1024,2048) -- weights regenerated from a seed, only outputs stored
    cls = load_class("train_sdxl_zh.py", "MLP")
    sd0 = O.adapter_init(1024, 1280, 1024, 2048, seed=0)
    xg = torch.randn(2, 77, 1024, generator=torch.Generator().manual_seed(1234))
    m, sd, outs, grads, dx = run_adapter(cls, (1024, 1280, 1024, 2048, False), xg, 0, True, sd_override=sd0)
    assert sum(v.numel() for v in sd.values()) == 6033408
    npz("adapter_sdxl_default", x1=outs["x1"], x2_sub=outs["x2"][:, ::7, ::16],
        x2_sum=outs["x2"].double().sum(), x2_sqsum=(outs["x2"].double() ** 2).sum(),
        keys=np.array(list(sd.keys())), shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
        **{"gradnorm." + k[5:]: v.double().norm() for k, v in grads.items()},
        grad_fc_bias=grads["grad.fc.bias"], grad_ln_weight=grads["grad.layernorm.weight"],
        grad_p0_sub=grads["grad.projector.0.weight"][::32, ::32])
    # shipped default A2 (1024,768,2048)
    cls = load_class("train_sd_zh.py", "MLP")
    sd0 = O.adapter_init(1024, 768, 2048, None, seed=0)
    m, sd, outs, grads, dx = run_adapter(cls, (1024, 768, 2048), xg, 0, False, sd_override=sd0)
    assert sum(v.numel() for v in sd.values()) == 7866368
    npz("adapter_sd_default", x2_sub=outs["x2"][:, ::7, ::8], x2_sum=outs["x2"].double().sum(),
        x2_sqsum=(outs["x2"].double() ** 2).sum(), keys=np.array(list(sd.keys())),
        grad_p4_sub=grads["grad.projector.4.weight"][::16, ::32])


def ref_lines(relpath, lo, hi, drop=("self.log(", "lr = self.trainer", "print(")):
    lines = open(os.path.join(REF, relpath)).read().split("\n")[lo - 1:hi]
    keep = []
    for ln in lines:
        if any(d in ln for d in drop):
            # keep block structure valid: a dropped print() inside an else: needs a body
            keep.append(ln[: len(ln) - len(ln.lstrip())] + "pass")
        else:
            keep.append(ln)
    return textwrap.dedent("\n".join(keep))


def run_ref_loss(script, base_rng, teacher_rng, num_blocks, tensors):
    """Execute the reference's own loss statements on the given tensors; returns (total, l_base, l_teacher, l_feat)."""
    args = types.SimpleNamespace(KD=True, hybrid_training=True)
    self_ = types.SimpleNamespace(KD_teacher=tensors["taps_t"], KD_student=tensors["taps_s"])
    ns = dict(torch=torch, F=F, args=args, self=self_, NUM_blocks=num_blocks, batch={"zh_or_not": tensors["zh"]},
              noise_pred=tensors["pred"], noise=tensors["noise"], noise_pred_teacher=tensors["pred_t"])
    exec(ref_lines(script, *base_rng), ns)
    l_base = ns["loss"].detach().clone()
    exec(ref_lines(script, *teacher_rng), ns)
    return ns["loss"], l_base, ns["loss_teacher"].detach(), torch.as_tensor(ns["loss_features"]).detach()


def losses():
    g = torch.Generator().manual_seed(7)
    B = 4
    zh = torch.tensor([0, 1, 1, 0])
    # SDXL: 7 taps (train_sdxl_zh.py:400-405 base, :416-441 KD terms)
    shapes = {"d0": (B, 8, 6, 6), "d1": (B, 16, 3, 3), "d2": (B, 32, 3, 3), "m": (B, 32, 3, 3),
              "u0": (B, 32, 6, 6), "u1": (B, 16, 12, 12), "u2": (B, 8, 12, 12)}
    t = {"pred": torch.randn(B, 4, 12, 12, generator=g).requires_grad_(True), "noise": torch.randn(B, 4, 12, 12, generator=g),
         "pred_t": torch.randn(B, 4, 12, 12, generator=g), "zh": zh,
         "taps_s": {k: torch.randn(s, generator=g).requires_grad_(True) for k, s in shapes.items()},
         "taps_t": {k: torch.randn(s, generator=g) for k, s in shapes.items()}}
    total, l0, l1, l2 = run_ref_loss("train_sdxl_zh.py", (400, 405), (416, 441), 3, t)
    total.backward()
    npz("loss_sdxl", zh=zh, pred=t["pred"], noise=t["noise"], pred_t=t["pred_t"], total=total, l0=l0, l1=l1, l2=l2,
        dpred=t["pred"].grad, order=np.array(list(shapes.keys())),
        **{"s." + k: v for k, v in t["taps_s"].items()}, **{"t." + k: v for k, v in t["taps_t"].items()},
        **{"ds." + k: v.grad for k, v in t["taps_s"].items()})
    # SD1.5: 9 taps with NaN/Inf skipping (train_sd_zh.py:218-223 base, :232-272 KD terms); tap u1 of the teacher
    # carries an Inf so that term must be skipped.
    shapes = {"d0": (B, 8, 6, 6), "d1": (B, 16, 3, 3), "d2": (B, 32, 2, 2), "d3": (B, 32, 1, 1), "m": (B, 32, 1, 1),
              "u0": (B, 32, 2, 2), "u1": (B, 32, 3, 3), "u2": (B, 16, 6, 6), "u3": (B, 8, 12, 12)}
    t = {"pred": torch.randn(B, 4, 12, 12, generator=g).requires_grad_(True), "noise": torch.randn(B, 4, 12, 12, generator=g),
         "pred_t": torch.randn(B, 4, 12, 12, generator=g), "zh": zh,
         "taps_s": {k: torch.randn(s, generator=g).requires_grad_(True) for k, s in shapes.items()},
         "taps_t": {k: torch.randn(s, generator=g) for k, s in shapes.items()}}
    t["taps_t"]["u1"][0, 3, 1, 1] = float("inf")
    total, l0, l1, l2 = run_ref_loss("train_sd_zh.py", (218, 223), (232, 272), 4, t)
    total.backward()
    zero = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in t["taps_s"].items()}
    npz("loss_sd_nanskip", zh=zh, pred=t["pred"], noise=t["noise"], pred_t=t["pred_t"], total=total, l0=l0, l1=l1, l2=l2,
        dpred=t["pred"].grad, order=np.array(list(shapes.keys())),
        **{"s." + k: v for k, v in t["taps_s"].items()}, **{"t." + k: v for k, v in t["taps_t"].items()},
        **{"ds." + k: v for k, v in zero.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not found; goldens can only be regenerated in the build container")
    torch.set_num_threads(4)
    adapters()
    losses()
