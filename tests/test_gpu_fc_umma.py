"""Parity test of the tcgen05 classifier kernels (csrc/dg_umma_fc.cu, dg_set_tuning key 14, on by default since round 2).
It runs the fused critic iteration and the generator iteration with key 14 off (CUDA-core classifier) and on at cfg-1 /
cfg-2 / ragged batch sizes and compares scalars and gradients (bf16 operand rounding of dz / classifier.0.weight is the only
intended difference); the absolute check against the oracle is tests/test_gpu_masks.py, which runs with the default (on)."""
import pytest
import torch

from downgan_b200 import _lib
from downgan_b200.synthetic import synth_batch
from oracle import networks as onet

import parity_util as pu
from test_gpu_parity import _run_steps

pytestmark = pytest.mark.gpu


def _flat(d):
    return torch.cat([v.reshape(-1).double() for v in d.values()])


@pytest.mark.parametrize("batch", [16, 64, 5])
def test_fc_umma_matches_cuda_core_path(batch):
    lib = _lib.load()
    G, C, _, _ = pu.build_pair(onet.GeneratorSpec(filters=16, channels=2), onet.CriticSpec(coarse_dim=16, fine_dim=128, nc=2),
                               "bf16", seed=0, critic_scale=1.9)
    coarse, fine, alpha = synth_batch(batch, 2, 16, seed=11, aseed=12)
    prev = lib.dg_set_tuning(14, 0)
    try:
        ref = _run_steps(G, C, coarse, fine, alpha)
        lib.dg_set_tuning(14, 1)
        got = _run_steps(G, C, coarse, fine, alpha)
    finally:
        lib.dg_set_tuning(14, prev)
    # scalars: critic loss, means, gp; generator loss
    assert pu.rel(got[0], ref[0]) < 5e-3 and pu.rel(got[2], ref[2]) < 5e-3
    cg, cr = got[1], ref[1]
    for k in cr:
        assert pu.rel(cg[k], cr[k]) < 3e-2, k   # bf16 rounding of dz and of classifier.0.weight
    assert float((_flat(got[3]) - _flat(ref[3])).norm() / _flat(ref[3]).norm()) < 3e-2
