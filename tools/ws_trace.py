"""Timeline of the TMA-fed conv kernel's three roles (DG_WS_TRACE=1: %globaltimer stamps of producer / MMA / epilogue per tile) on
the late critic layers.  Usage on the GPU box:  DG_WS_TRACE=1 python tools/ws_trace.py 2> trace.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu

cases = [(192, 64, 128, 16, 1), (192, 128, 128, 16, 2), (192, 64, 64, 32, 2), (192, 32, 64, 32, 1), (192, 16, 32, 64, 1)]
for (b, ci, co, h, s) in cases:
    g = torch.Generator().manual_seed(7)
    x = torch.randn(b, ci, h, h, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** .5)
    bias = torch.randn(co, generator=g)
    ho = h // s
    dy = torch.randn(b, co, ho, ho, generator=g)
    for rep in range(2):
        print(f"=== b{b} ci{ci} co{co} {h}x{h} s{s} fwd rep{rep}", file=sys.stderr, flush=True)
        pu.conv_fwd(x, wt, bias, s, 0.2, "bf16")
    print(f"=== b{b} ci{ci} co{co} {h}x{h} s{s} dgrad", file=sys.stderr, flush=True)
    pu.conv_dgrad(dy, wt, h, h, s, "bf16")
    pu.conv_dgrad(dy, wt, h, h, s, "bf16")
