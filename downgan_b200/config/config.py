"""Run-time device selection, mirroring ``DoWnGAN/config/config.py:25``.  Under
``torchrun`` every rank uses its own ``LOCAL_RANK`` GPU."""
import os

import torch

device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
