"""Diagnostic: tcgen05 conv kernel (bf16 mode primitives) vs torch conv on bf16-rounded operands."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import parity_util as pu
from downgan_b200 import _lib

def r16(t):
    return t.bfloat16().float()

def main():
    lib = _lib.load()
    print("tcgen05 compiled:", lib.dg_has_tcgen05())
    cases = [(1, 16, 16, 16, 16), (3, 16, 16, 16, 16), (2, 80, 16, 16, 16), (2, 32, 16, 16, 16), (2, 16, 64, 32, 32),
             (1, 48, 16, 8, 8), (2, 16, 32, 64, 64), (2, 32, 64, 32, 32), (1, 16, 16, 128, 128), (2, 64, 32, 32, 32),
             (1, 64, 128, 16, 16), (2, 16, 16, 20, 12)]
    for (b, ci, co, h, w) in cases:
        g = torch.Generator().manual_seed(7)
        x = r16(torch.randn(b, ci, h, w, generator=g)); wt = r16(torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5))
        bias = torch.randn(co, generator=g)
        y_ref = F.leaky_relu(F.conv2d(x, wt, bias, padding=1), 0.2)
        lib.dg_profile(1)
        y = pu.conv_fwd(x, wt, bias, 1, 0.2, "bf16")
        dy = r16(torch.randn_like(y_ref))
        dx = pu.conv_dgrad(dy, wt, h, w, 1, "bf16")
        import ctypes as C
        buf = (C.c_double * 44)()
        lib.dg_profile_report(buf, 11); lib.dg_profile(0)
        dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, padding=1)
        e = (y - y_ref).abs()
        print(f"case b{b} ci{ci} co{co} {h}x{w}: fwd rel {pu.rel(y, y_ref):.2e} maxabs {float(e.max()):.3e} | dgrad rel {pu.rel(dx, dx_ref):.2e} | "
              f"umma launches {int(buf[8])} direct {int(buf[0])}")
        if pu.rel(y, y_ref) > 1e-2:
            bad = (e > 0.05).nonzero()
            print("   first bad idx:", bad[:6].tolist(), "of", bad.shape[0], "/", y.numel())

if __name__ == "__main__":
    main()
