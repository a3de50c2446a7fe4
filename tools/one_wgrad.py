"""Run one weight gradient through the bf16 primitive (for ncu captures): b ci co h stride."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
b, ci, co, h, s = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (64, 2, 16, 128, 1))]
x = torch.randn(b, ci, h, h); ho = (h - 1) // s + 1; dy = torch.randn(b, co, ho, ho)
for _ in range(3):
    dw, db = pu.conv_wgrad(x, dy, s, "bf16")
torch.cuda.synchronize()
print("ok", dw.shape)
os._exit(0)
