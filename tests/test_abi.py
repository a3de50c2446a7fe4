"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares, the flat parameter layout agrees with the reference state_dict, and
the host modules keep the reference interface.  No compute calls (no GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

from downgan_b200 import _lib
from downgan_b200.networks import Critic, Generator
from oracle import networks as onet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "downgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/downgan_b200.h but not exported"
    assert set(syms) == set(_lib.EXPORTS), set(syms) ^ set(_lib.EXPORTS)
    assert lib.dg_abi_version() == 2


@pytest.mark.parametrize("filters,channels,rrdb", [(16, 2, 16), (16, 7, 16), (32, 7, 23), (8, 3, 2)])
def test_generator_layout_matches_state_dict(lib, filters, channels, rrdb):
    torch.manual_seed(0)
    g = Generator(filters, filters * 8, channels, 2, num_res_blocks=rrdb)
    spec = onet.GeneratorSpec(filters, channels, 2, rrdb, 3)
    keys = onet.generator_keys(spec)
    sd = g.state_dict()
    assert list(sd.keys()) == [k for k, _ in keys]
    assert [n for n, _ in g.named_parameters()] == [k for k, _ in keys]
    cfg = _lib.GeneratorConfig(filters, channels, 2, rrdb, 3, filters, 1, 0)
    off = 0
    for i, (k, shp) in enumerate(keys):
        assert tuple(sd[k].shape) == shp and sd[k].dtype == torch.float32
        assert lib.dg_generator_param_offset(C.byref(cfg), i) == off, k
        off += sd[k].numel()
    assert lib.dg_generator_param_count(C.byref(cfg)) == off
    assert lib.dg_generator_param_offset(C.byref(cfg), len(keys)) == -1


@pytest.mark.parametrize("w,fine", [(16, 128), (32, 256), (8, 64)])
def test_critic_layout_matches_state_dict(lib, w, fine):
    torch.manual_seed(0)
    c = Critic(w, fine, 2)
    spec = onet.CriticSpec(w, fine, 2)
    keys = onet.critic_keys(spec)
    sd = c.state_dict()
    assert list(sd.keys()) == [k for k, _ in keys]
    cfg = _lib.CriticConfig(w, fine, 2, 1, 0)
    off = 0
    for i, (k, shp) in enumerate(keys):
        assert tuple(sd[k].shape) == shp
        assert lib.dg_critic_param_offset(C.byref(cfg), i) == off, k
        off += sd[k].numel()
    assert lib.dg_critic_param_count(C.byref(cfg)) == off


def test_default_init_equals_reference_stream():
    """Same seed -> same weights as the reference constructors (pinned in oracle/make_golden.py)."""
    torch.manual_seed(0)
    c = Critic(8, 64, 2)
    g = Generator(8, 64, 3, 2, num_res_blocks=2)
    torch.manual_seed(0)
    c_sd = onet.init_critic_state(onet.CriticSpec(8, 64, 2))
    g_sd = onet.init_generator_state(onet.GeneratorSpec(8, 3, 2, 2, 3))
    for k, v in c.state_dict().items():
        assert torch.equal(v, c_sd[k]), k
    for k, v in g.state_dict().items():
        assert torch.equal(v, g_sd[k]), k


def test_load_state_dict_roundtrip():
    torch.manual_seed(1)
    g_sd = onet.init_generator_state(onet.GeneratorSpec(8, 3, 2, 2, 3))
    g = Generator(8, 64, 3, 2, num_res_blocks=2)
    g.load_state_dict(g_sd)
    for k, v in g.state_dict().items():
        assert torch.equal(v, g_sd[k])


def test_no_cpu_fallback():
    g = Generator(8, 64, 3, 2, num_res_blocks=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        g(torch.zeros(1, 3, 8, 8))


def test_bad_arguments_report_errors(lib):
    h = C.c_void_p()
    cfg = _lib.CriticConfig(16, 100, 2, 1, 0)  # fine_dim not a multiple of 16
    assert lib.dg_critic_create(C.byref(cfg), C.byref(h)) == -1
    assert b"fine_dim" in lib.dg_last_error()
    assert lib.dg_adam_step(None, None, None, None, 0, 0.1, 0.9, 0.99, 1e-8, 1, 1.0, None) == -1


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under downgan_b200/ may reference it."""
    bad = []
    for dp, _dn, fn in os.walk(os.path.join(ROOT, "downgan_b200")):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_sass_descriptor_pairs_are_written():
    """tools/sass_desc_check.py on every built object: each tcgen05.mma descriptor operand is a 64-bit uniform register pair, and
    both halves must be written by the kernel's own arithmetic.  (Round 2: ptxas dropped the high-word update of one pair in an
    unrolled loop of fc_wgrad_umma_kernel; the MMA then read a stale parameter word as its stride - an illegal shared-memory
    access that came and went with the address of an input buffer.)"""
    import glob
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    objs = sorted(glob.glob(os.path.join(root, "downgan_b200", "csrc", "*.o")))
    if not objs:
        pytest.skip("no object files (library built elsewhere)")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_desc_check.py"), *objs], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
