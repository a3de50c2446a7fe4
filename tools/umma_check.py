"""Diagnostic: tcgen05 kernels (bf16-mode primitives) vs torch on bf16-rounded operands."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import parity_util as pu
from downgan_b200 import _lib

def r16(t):
    return t.bfloat16().float()

def counts(lib):
    buf = (C.c_double * 44)()
    lib.dg_profile_report(buf, 11)
    return {n: int(buf[4 * i]) for i, n in enumerate(_lib.PROFILE_CLASSES) if buf[4 * i] > 0}

def main():
    lib = _lib.load()
    print("tcgen05 compiled:", lib.dg_has_tcgen05())
    cases = [(1, 16, 16, 16, 16, 1), (3, 16, 16, 16, 16, 1), (2, 80, 16, 16, 16, 1), (2, 16, 64, 32, 32, 1),
             (1, 48, 16, 8, 8, 1), (2, 16, 32, 64, 64, 1), (2, 32, 64, 32, 32, 1), (1, 16, 16, 128, 128, 1),
             (1, 64, 128, 16, 16, 1), (2, 16, 16, 20, 12, 1), (3, 128, 64, 16, 16, 1),
             (2, 16, 16, 64, 64, 2), (2, 32, 32, 32, 32, 2), (3, 64, 64, 16, 16, 2), (2, 128, 128, 16, 16, 2),
             (2, 16, 16, 128, 128, 2), (5, 48, 96, 8, 8, 1)]
    for (b, ci, co, h, w, s) in cases:
        g = torch.Generator().manual_seed(7)
        x = r16(torch.randn(b, ci, h, w, generator=g)); wt = r16(torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5))
        bias = torch.randn(co, generator=g)
        y_ref = F.leaky_relu(F.conv2d(x, wt, bias, stride=s, padding=1), 0.2)
        dy = r16(torch.randn_like(y_ref))
        lib.dg_profile(1)
        y = pu.conv_fwd(x, wt, bias, s, 0.2, "bf16")
        dx = pu.conv_dgrad(dy, wt, h, w, s, "bf16")
        dw, dbias = pu.conv_wgrad(x, dy, s, "bf16")
        cnt = counts(lib); lib.dg_profile(0)
        dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=1)
        dw_ref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=s, padding=1)
        print(f"case b{b} ci{ci} co{co} {h}x{w} s{s}: fwd {pu.rel(y, y_ref):.2e} dgrad {pu.rel(dx, dx_ref):.2e} "
              f"wgrad {pu.rel(dw, dw_ref):.2e} dbias {pu.rel(dbias, dy.sum((0,2,3))):.2e} | {cnt}")
        if pu.rel(dw, dw_ref) > 1e-2:
            e = (dw - dw_ref).abs()
            for tap in range(9):
                print("    tap", tap, "rel", f"{pu.rel(dw[:, :, tap // 3, tap % 3], dw_ref[:, :, tap // 3, tap % 3]):.2e}")

if __name__ == "__main__":
    main()
