"""Host logic of the HBM-resident loader (CPU): its epoch permutations must be the index stream torch's own
DataLoader(shuffle=True) — what the reference builds in GAN/stage.py:73-81 — produces for the same seed, epoch after epoch,
including the reference's two `next(iter(dataloader))` calls per epoch (wasserstein.py:155,175)."""
import pytest
import torch

from downgan_b200.GAN.dataloader import DeviceLoader, NetCDFSR, epoch_permutation
from oracle import dataloader as odl


def _ids(n):
    # sample i carries its own index in both tensors, so a batch reveals which rows it holds
    coarse = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1, 1).expand(n, 2, 2, 2).contiguous()
    fine = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1, 1).expand(n, 2, 4, 4).contiguous()
    return coarse, fine


@pytest.mark.parametrize("n,bs", [(37, 8), (64, 32), (5, 7)])
def test_permutation_stream_equals_torch_dataloader(n, bs):
    coarse, fine = _ids(n)
    torch.manual_seed(123)
    ref = odl.reference_loader(coarse, fine, bs, shuffle=True)
    ref_epochs = []
    for _ in range(3):
        ref_epochs.append([c[:, 0, 0, 0].long() for c, _f in ref])
        next(iter(ref))  # the reference's plot batches (train set)
    torch.manual_seed(123)
    dl = DeviceLoader(NetCDFSR(coarse, fine), bs, shuffle=True)
    assert len(dl) == len(ref)
    for e in range(3):
        got = dl.batch_indices()
        assert len(got) == len(ref_epochs[e])
        for a, b in zip(got, ref_epochs[e]):
            assert torch.equal(a, b)
        dl.skip_first_batch()


def test_sequential_and_drop_last():
    coarse, fine = _ids(10)
    dl = DeviceLoader(NetCDFSR(coarse, fine), 4, shuffle=False)
    assert [b.tolist() for b in dl.batch_indices()] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    dl = DeviceLoader(NetCDFSR(coarse, fine), 4, shuffle=False, drop_last=True)
    assert len(dl) == 2 and [b.tolist() for b in dl.batch_indices()] == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_explicit_generator_matches_torch():
    coarse, fine = _ids(23)
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    from torch.utils.data import DataLoader
    ref = DataLoader(odl.NetCDFSR(coarse, fine), batch_size=5, shuffle=True, generator=g1)
    want = [c[:, 0, 0, 0].long() for c, _ in ref]
    got = DeviceLoader(NetCDFSR(coarse, fine), 5, shuffle=True, generator=g2).batch_indices()
    assert all(torch.equal(a, b) for a, b in zip(got, want))


def test_dataset_protocol_and_cpu_refusal():
    coarse, fine = _ids(6)
    ds = NetCDFSR(coarse, fine, device=torch.device("cpu"))
    assert len(ds) == 6
    c, f = ds[torch.tensor(3)]
    assert float(c[0, 0, 0]) == 3.0 and f.shape == (2, 4, 4)
    with pytest.raises(ValueError):
        NetCDFSR(coarse, fine[:5])
    with pytest.raises(Exception):  # no CPU fallback: the gather is a CUDA kernel
        next(iter(DeviceLoader(ds, 2)))
    assert epoch_permutation(4, shuffle=False).tolist() == [0, 1, 2, 3]
