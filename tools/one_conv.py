"""Run one critic-like conv through the bf16 primitive (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
b, ci, co, h, s = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (96, 16, 16, 128, 2))]
x = torch.randn(b, ci, h, h); w = torch.randn(co, ci, 3, 3) / (3 * ci ** .5)
for _ in range(3):
    y = pu.conv_fwd(x, w, None, s, 0.2, "bf16")
torch.cuda.synchronize()
print("ok", y.shape)
os._exit(0)
