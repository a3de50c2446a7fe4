"""A/B of the two tcgen05 conv kernels per layer shape: the streaming implicit-GEMM kernel (dg_umma_conv_ig.cu) against the
weights-stationary kernel (dg_umma_conv_ws.cu), kernel time from the library's CUDA-event profiler (best of 5), plus the
relative error of both against a torch fp32 conv.  Run on the GPU box with DG_IG=2 (IG wherever supported; dg_set_tuning(16, 0)
switches it off at run time):   DG_IG=2 python tools/ig_ab.py > gpurun_out/ig_ab.md"""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import parity_util as pu
from downgan_b200 import _lib


def prof_ms(lib, cls="conv_tcgen05"):
    buf = (C.c_double * 44)()
    lib.dg_profile_report(buf, 11)
    i = _lib.PROFILE_CLASSES.index(cls)
    return buf[4 * i], buf[4 * i + 1]


def timed(lib, fn, reps=5):
    out = fn()
    best = 1e9
    for _ in range(reps):
        lib.dg_profile(1)
        fn()
        n, t = prof_ms(lib)
        lib.dg_profile(0)
        if n > 0:
            best = min(best, 1e3 * t / n)
    return best, out


def main():
    lib = _lib.load()
    # (name, b, ci, co, h_in, stride): critic layers 1..7 at 3B = 192 and B = 64, dense-block shapes of cfg-4 / cfg-5
    cases = [("L7", 192, 128, 128, 16, 2), ("L6", 192, 64, 128, 16, 1), ("L5", 192, 64, 64, 32, 2), ("L4", 192, 32, 64, 32, 1),
             ("L3", 192, 32, 32, 64, 2), ("L2", 192, 16, 32, 64, 1), ("L1", 192, 16, 16, 128, 2),
             ("L7 B64", 64, 128, 128, 16, 2), ("L6 B64", 64, 64, 128, 16, 1), ("L5 B64", 64, 64, 64, 32, 2), ("L4 B64", 64, 32, 64, 32, 1),
             ("cfg4 db k5", 32, 160, 32, 32, 1), ("cfg4 db k3", 32, 96, 32, 32, 1), ("cfg5 db k5", 16, 320, 64, 64, 1),
             ("cfg5 db k3", 16, 192, 64, 64, 1), ("cfg4 C L7", 96, 256, 256, 32, 2)]
    print("| layer | shape (B, Ci->Co, HxH in, stride) | pass | ws µs | ig µs | ws/ig | ws err | ig err |")
    print("|---|---|---|---|---|---|---|---|")
    for (name, b, ci, co, h, s) in cases:
        g = torch.Generator().manual_seed(7)
        x = torch.randn(b, ci, h, h, generator=g)
        wt = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5)
        bias = torch.randn(co, generator=g)
        ho = h // s
        dy = torch.randn(b, co, ho, ho, generator=g)
        y_ref = F.leaky_relu(F.conv2d(x, wt, bias, stride=s, padding=1), 0.2)
        dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=1)
        for what, fn, ref in (("fwd", lambda: pu.conv_fwd(x, wt, bias, s, 0.2, "bf16"), y_ref),
                              ("dgrad", lambda: pu.conv_dgrad(dy, wt, h, h, s, "bf16"), dx_ref)):
            lib.dg_set_tuning(16, 0)
            t_ws, o_ws = timed(lib, fn)
            lib.dg_set_tuning(16, 1)
            t_ig, o_ig = timed(lib, fn)
            print(f"| {name} | {b}, {ci}->{co}, {h}x{h}, s{s} | {what} | {t_ws:.1f} | {t_ig:.1f} | {t_ws / t_ig:.2f} | "
                  f"{pu.rel(o_ws, ref):.2e} | {pu.rel(o_ig, ref):.2e} |", flush=True)


if __name__ == "__main__":
    main()
