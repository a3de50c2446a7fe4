"""Diagnostic: A/B of the tcgen05 conv kernels (dg_set_tuning key 0) on the cfg-2 layer shapes:
correctness against torch on bf16-rounded operands and the kernel time from the library's own
CUDA-event profiler (class conv_tcgen05)."""
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


def prof_ms(lib, cls="conv_tcgen05"):
    buf = (C.c_double * 44)()
    lib.dg_profile_report(buf, 11)
    i = _lib.PROFILE_CLASSES.index(cls)
    return buf[4 * i], buf[4 * i + 1]


def main():
    lib = _lib.load()
    # (b, ci, co, h, stride): critic layers 2..8 at 3B = 192 and generator tail at B = 64
    cases = [(192, 16, 16, 128, 2), (192, 16, 32, 64, 1), (192, 32, 32, 64, 2), (192, 32, 64, 32, 1),
             (192, 64, 64, 32, 2), (192, 64, 128, 16, 1), (192, 128, 128, 16, 2),
             (64, 16, 16, 128, 1), (64, 16, 64, 64, 1), (64, 16, 16, 16, 1), (3, 16, 16, 20, 1), (2, 32, 16, 12, 2)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = [(4, ci, co, h, s) for (_, ci, co, h, s) in cases]
    if "carve" in sys.argv:
        print("carveout: prefer shared", lib.dg_set_tuning(1, 1))
    elif "nocarve" in sys.argv:
        print("carveout: no preference", lib.dg_set_tuning(1, 0))
    for (b, ci, co, h, s) in cases:
        g = torch.Generator().manual_seed(7)
        x = r16(torch.randn(b, ci, h, h, generator=g))
        wt = r16(torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5))
        bias = torch.randn(co, generator=g)
        bref = min(b, 8)
        y_ref = F.leaky_relu(F.conv2d(x[:bref], wt, bias, stride=s, padding=1), 0.2)
        dy = r16(torch.randn(b, co, y_ref.shape[2], y_ref.shape[3], generator=g))
        dx_ref = torch.nn.grad.conv2d_input(x[:bref].shape, wt, dy[:bref], stride=s, padding=1)
        line = f"b{b} ci{ci} co{co} {h}x{h} s{s}:"
        for ws in (1, 0):
            lib.dg_set_tuning(0, ws)
            y = pu.conv_fwd(x, wt, bias, s, 0.2, "bf16")  # warm
            lib.dg_profile(1)
            y = pu.conv_fwd(x, wt, bias, s, 0.2, "bf16")
            n1, t1 = prof_ms(lib); lib.dg_profile(0)
            dx = pu.conv_dgrad(dy, wt, h, h, s, "bf16")
            lib.dg_profile(1)
            dx = pu.conv_dgrad(dy, wt, h, h, s, "bf16")
            n2, t2 = prof_ms(lib); lib.dg_profile(0)
            line += (f"  [{'ws ' if ws else 'old'}] fwd {1e3 * t1:7.1f}us err {pu.rel(y[:bref], y_ref):.1e}"
                     f" | dgrad {1e3 * t2:7.1f}us err {pu.rel(dx[:bref], dx_ref):.1e}")
        lib.dg_set_tuning(0, 1)
        print(line, flush=True)


if __name__ == "__main__":
    main()
