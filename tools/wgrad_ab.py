"""Diagnostic: A/B of the tcgen05 weight-gradient kernels (dg_set_tuning key 2) on the cfg-2 layer shapes."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
from downgan_b200 import _lib


def r16(t):
    return t.bfloat16().float()


def prof_ms(lib, cls="wgrad_tcgen05"):
    buf = (C.c_double * 44)()
    lib.dg_profile_report(buf, 11)
    i = _lib.PROFILE_CLASSES.index(cls)
    return buf[4 * i], buf[4 * i + 1]


def main():
    lib = _lib.load()
    cases = [(192, 2, 16, 128, 1), (64, 2, 16, 128, 1), (192, 16, 16, 128, 2), (192, 16, 32, 64, 1), (192, 32, 32, 64, 2), (192, 32, 64, 32, 1),
             (192, 64, 64, 32, 2), (192, 64, 128, 16, 1), (192, 128, 128, 16, 2),
             (64, 16, 16, 128, 1), (64, 16, 64, 64, 1), (64, 16, 64, 16, 1), (3, 16, 16, 20, 1), (2, 32, 16, 12, 2), (5, 64, 32, 10, 1)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = [(4, ci, co, h, s) for (_, ci, co, h, s) in cases]
    for (b, ci, co, h, s) in cases:
        g = torch.Generator().manual_seed(7)
        x = r16(torch.randn(b, ci, h, h, generator=g))
        ho = (h - 1) // s + 1
        dy = r16(torch.randn(b, co, ho, ho, generator=g))
        bref = min(b, 6)
        dw_ref = torch.nn.grad.conv2d_weight(x[:bref].double(), (co, ci, 3, 3), dy[:bref].double(), stride=s, padding=1).float()
        line = f"b{b} ci{ci} co{co} {h}x{h} s{s}:"
        for ws in (1, 0):
            lib.dg_set_tuning(2, ws)
            dw_small, _ = pu.conv_wgrad(x[:bref], dy[:bref], s, "bf16")
            dw, db = pu.conv_wgrad(x, dy, s, "bf16")  # warm
            lib.dg_profile(1)
            dw, db = pu.conv_wgrad(x, dy, s, "bf16")
            n1, t1 = prof_ms(lib); lib.dg_profile(0)
            line += f"  [{'ws ' if ws else 'old'}] {1e3 * t1:7.1f}us ({int(n1)} launches) err {pu.rel(dw_small, dw_ref):.1e} dbias {pu.rel(db, dy.sum((0, 2, 3))):.1e}"
        lib.dg_set_tuning(2, 1)
        print(line, flush=True)


if __name__ == "__main__":
    main()
