"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares, the product
refuses host tensors (no CPU fallback), checkpoint-key compatibility, and the data-parallel gradient exchange on a
world_size-2 gloo group."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pea_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pea_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from pea_diffusion_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 15
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/pea_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding table and header disagree"
    assert _lib.lib.pea_version() == 201
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "pea_gemm" in nm


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """Every descriptor struct of include/pea_b200.h against its ctypes mirror in _lib.py: total size and the offset of every
    field, taken from a C program compiled against the header (the way a maintainer binding from C / cgo would see it)."""
    import ctypes
    from pea_diffusion_b200 import _lib
    mirrors = {"pea_kd_tap": _lib.KdTap, "pea_attn2_desc": _lib.Attn2Desc, "pea_gemm_desc": _lib.GemmDesc,
               "pea_split_item": _lib.SplitItem}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "pea_b200.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        cls = mirrors[cname]
        want = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == want, f"{cname}.{field}: header {value}, ctypes {want}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())


def test_plain_c_caller_links_and_gets_errors_as_text(tmp_path):
    """examples/abi_from_c.c: the header compiles as C99, the library links from plain C, version matches, malformed calls
    return PEA_E* codes with a message (no GPU needed for any of that)."""
    from pea_diffusion_b200 import _lib
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "abi_from_c"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "abi_from_c.c"),
                    "-L", libdir, "-lpea_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pea_attn2_fwd with null pointers -> -1: attn2: null" in r.stdout and "pea_gemm with a bad descriptor -> -1" in r.stdout


def test_argument_errors_come_back_as_text_without_a_gpu():
    import ctypes as C
    from pea_diffusion_b200 import _lib
    d = _lib.GemmDesc()
    assert _lib.lib.pea_gemm(C.byref(d), None) == -1
    assert b"M, N, K" in _lib.lib.pea_last_error()
    a = _lib.Attn2Desc()
    a.B = a.N = a.H = 1
    a.L = 100
    assert _lib.lib.pea_attn2_fwd(C.byref(a), None) == -1 and b"L=100" in _lib.lib.pea_last_error()


def test_product_refuses_host_tensors():
    from pea_diffusion_b200 import MLP, PEAAttn2Processor, kd_loss
    from pea_diffusion_b200._lib import PeaError
    from pea_diffusion_b200.shell.unet import Attention
    with pytest.raises(PeaError, match="no CPU fallback"):
        MLP(64, 80, 96, 128, False)(torch.randn(2, 7, 64))
    attn = Attention(64, 128, 1, 64).half()
    with pytest.raises(PeaError, match="no CPU fallback"):
        PEAAttn2Processor()(attn, torch.randn(1, 16, 64).half(), encoder_hidden_states=torch.randn(1, 77, 128).half())
    p = torch.randn(2, 4, 8, 8)
    with pytest.raises(Exception):
        kd_loss(p, p, p, [], [], torch.tensor([0, 1]))


def test_state_dict_keys_match_reference_goldens():
    from pea_diffusion_b200 import MLP, MLP_Legacy, MLP_SD
    from tests.util import load_golden, state_dict_of
    for name, mod in (("adapter_sdxl_small", MLP(64, 80, 96, 128, False)), ("adapter_sd_small", MLP_SD(64, 48, 96)),
                      ("adapter_sd5_small", MLP_SD(64, 48, 96, num_linear=5)), ("adapter_legacy_small", MLP_Legacy(64, 80, 96, 128, False))):
        sd = state_dict_of(load_golden(name))
        assert list(mod.state_dict().keys()) == list(sd.keys()), name
        mod.load_state_dict(sd, strict=True)
    full = MLP(1024, 1280, 1024, 2048, False)
    assert sum(p.numel() for p in full.parameters()) == 6033408  # README.md:11 "only 6M parameters"
    with pytest.raises(AssertionError):
        MLP(64, 80, 96, 128, True)  # use_residual needs in_dim == out_dim (train_sdxl_zh.py:45-46)


def test_processor_falls_back_outside_contract():
    """Self-attention / masks / fp32 activations are routed to the stock processor."""
    from pea_diffusion_b200 import PEAAttn2Processor
    from pea_diffusion_b200.shell.unet import Attention, StockAttnProcessor
    torch.manual_seed(0)
    attn = Attention(64, None, 2, 32)
    x = torch.randn(1, 16, 64)
    want = StockAttnProcessor()(attn, x)
    got = PEAAttn2Processor(fallback=StockAttnProcessor())(attn, x)
    assert torch.equal(want, got)


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pea_diffusion_b200.train import GradAllReducer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
m = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
red = GradAllReducer(m.parameters())
opt = torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.0)
g = torch.Generator().manual_seed(100 + rank)
full = [torch.randn(6, 8, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
for step in range(3):
    opt.zero_grad(set_to_none=True)
    m(full[rank]).pow(2).mean().backward()
    red.finish()
    opt.step()
# single-process replay on the concatenated batch must give the same parameters
torch.manual_seed(0)
ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.0)
for step in range(3):
    ropt.zero_grad(set_to_none=True)
    ref(torch.cat(full)).pow(2).mean().backward()
    ropt.step()
for p, q in zip(m.parameters(), ref.parameters()):
    assert torch.allclose(p, q, atol=1e-6), (p - q).abs().max()
flat = torch.cat([p.detach().flatten() for p in m.parameters()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(gathered[0], t) for t in gathered), "ranks diverged"
assert red.bytes_per_step == sum(p.numel() * 4 for p in m.parameters())
dist.destroy_process_group()
print("ok", rank)
'''


def test_grad_allreduce_world2_gloo(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on the tiny shell: exactly one JSON line on
    stdout with the contract's keys, whatever else the libraries print."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--tiny", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "SDXL-1024 KD train samples/s" and d["unit"] == "samples/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True
    # the reference arm must not map any of this repo's native code into its process
    assert d["native_libs_mapped"] == [], d["native_libs_mapped"]


def test_shell_and_oracle_import_without_binding_the_library():
    """Only the product modules bind libpea_b200.so (and fail loudly without it); the plain-PyTorch UNet shell that hosts
    the path and the oracle stay importable on their own -- that is what keeps the CPU reference arm free of our kernels."""
    code = ("import sys; sys.path.insert(0, %r); import pea_diffusion_b200.shell.unet, oracle.pea_oracle; "
            "assert 'pea_diffusion_b200._lib' not in sys.modules; "
            "assert not any('libpea_b200' in l for l in open('/proc/self/maps')); "
            "import pea_diffusion_b200 as p; p.MLP; assert 'pea_diffusion_b200._lib' in sys.modules; print('ok')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_batch_stager_needs_cuda():
    from pea_diffusion_b200 import BatchStager
    from pea_diffusion_b200._lib import PeaError
    with pytest.raises(PeaError, match="no CPU fallback"):
        BatchStager("cpu")


def test_attn2_group_layout_and_weight_cache_on_host_modules():
    """Attn2Group bookkeeping needs no GPU: column slices in forward-call order, one concatenated [Wk; Wv] per weight
    version, LoRA on any to_k / to_v sends the group back to the per-layer path."""
    from pea_diffusion_b200.attn2 import Attn2Group

    class A(torch.nn.Module):
        def __init__(self, c, dc):
            super().__init__()
            self.to_k = torch.nn.Linear(dc, c, bias=False)
            self.to_v = torch.nn.Linear(dc, c, bias=False)

    torch.manual_seed(0)
    attns = [A(64, 32), A(128, 32), A(64, 32)]
    g = Attn2Group(attns)
    assert g.widths == [128, 256, 128] and g.offsets == [0, 128, 384] and g.total == 512
    assert [g.index[id(a)] for a in attns] == [0, 1, 2]
    w = g.weights(torch.float32, torch.device("cpu"))
    assert w.shape == (512, 32) and g.weights(torch.float32, torch.device("cpu")) is w          # cached
    assert torch.equal(w[128:256], attns[1].to_k.weight) and torch.equal(w[256:384], attns[1].to_v.weight)
    with torch.no_grad():
        attns[2].to_v.weight.mul_(2.0)                                                           # version bump -> rebuilt
    w2 = g.weights(torch.float32, torch.device("cpu"))
    assert w2 is not w and torch.equal(w2[448:], attns[2].to_v.weight)
    assert not g.supported(torch.bfloat16)      # trainable to_k / to_v under autograd: the fused path forms no dWk / dWv
    for a in attns:
        a.requires_grad_(False)
    g._supported = None
    assert g.supported(torch.bfloat16)
    attns[0].to_k.lora_layer = object()
    g._supported = None
    assert not g.supported(torch.bfloat16)
