"""Sweep of the TMA-fed conv kernel's tiling (DG_WS_FORCE = NT:cps:TH) on chosen layer shapes: kernel time from the
library's CUDA-event profiler.  Usage: python tools/ws_sweep.py  (on the GPU box; plans go to stderr with DG_WS_PLAN=1)"""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
from downgan_b200 import _lib


def prof_ms(lib, cls="conv_tcgen05"):
    buf = (C.c_double * 44)()
    lib.dg_profile_report(buf, 11)
    i = _lib.PROFILE_CLASSES.index(cls)
    return buf[4 * i], buf[4 * i + 1]


def timed(lib, fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        lib.dg_profile(1)
        fn()
        n, t = prof_ms(lib)
        lib.dg_profile(0)
        if n > 0:
            best = min(best, 1e3 * t / n)
    return best


def main():
    lib = _lib.load()
    # (b, ci, co, h, stride): late critic layers at 3B = 192 and the JVP pass at B = 64
    cases = [(192, 64, 128, 16, 1), (192, 128, 128, 16, 2), (192, 64, 64, 32, 2), (192, 32, 64, 32, 1),
             (64, 64, 128, 16, 1), (64, 128, 128, 16, 2)]
    for (b, ci, co, h, s) in cases:
        g = torch.Generator().manual_seed(7)
        x = torch.randn(b, ci, h, h, generator=g)
        wt = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5)
        bias = torch.randn(co, generator=g)
        ho = h // s
        dy = torch.randn(b, co, ho, ho, generator=g)
        for what, fn, n_out in (("fwd", lambda: pu.conv_fwd(x, wt, bias, s, 0.2, "bf16"), co),
                                ("dgrad", lambda: pu.conv_dgrad(dy, wt, h, h, s, "bf16"), ci)):
            os.environ.pop("DG_WS_FORCE", None)
            res = [("auto", timed(lib, fn))]
            for nt in (16, 32, 64, 128):
                if nt > n_out:
                    continue
                for cps in (1, 2):
                    os.environ["DG_WS_FORCE"] = f"{nt}:{cps}:0"
                    res.append((f"{nt}/{cps}", timed(lib, fn)))
            os.environ.pop("DG_WS_FORCE", None)
            print(f"b{b} ci{ci} co{co} {h}x{h} s{s} {what:5s}: " + "  ".join(f"{k} {v:6.1f}" for k, v in res), flush=True)


if __name__ == "__main__":
    main()
