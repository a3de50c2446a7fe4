"""HBM-resident dataset and batch assembly: drop-in for ``DoWnGAN/GAN/dataloader.py:6-33`` (``NetCDFSR``) and the
``torch.utils.data.DataLoader(dataset, batch_size=hp.batch_size, shuffle=True)`` built in ``GAN/stage.py:73-81``.

The reference already keeps both tensors on the device (``stage.py:29-32``) and lets the DataLoader call
``__getitem__`` once per sample and ``torch.stack`` the results — ``batch_size`` indexing kernels plus a concatenation per
batch, driven from Python.  Here a batch is ONE row gather (``dg_gather_rows``: coalesced 16-byte loads, one launch per
tensor) driven by the epoch's permutation, which is uploaded once per epoch.

The permutation is drawn exactly as torch's ``RandomSampler`` draws it, consuming the global CPU RNG the same way
(``_BaseDataLoaderIter.__init__`` draws a base seed, ``RandomSampler.__iter__`` draws its own seed and calls
``torch.randperm(n, generator)``), so for the same ``torch.manual_seed`` the batches are the rows the reference's loader
would deliver, in the same order (``tests/test_dataloader.py`` checks this against torch's DataLoader itself).
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Tuple

import torch
from torch.utils.data import Dataset

from .. import _lib


class NetCDFSR(Dataset):
    """Data loader from torch.Tensors (same constructor and item protocol as the reference class)."""

    def __init__(self, coarse: torch.Tensor, fine: torch.Tensor, device: Optional[torch.device] = None) -> None:
        if coarse.shape[0] != fine.shape[0]:
            raise ValueError(f"coarse holds {coarse.shape[0]} samples, fine {fine.shape[0]}")
        self.fine = fine
        self.coarse = coarse
        self.device = device

    def __len__(self) -> int:
        return self.fine.size(0)

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return self.coarse[idx, ...], self.fine[idx, ...]


def epoch_permutation(n: int, shuffle: bool = True, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Row order of one epoch, consuming RNG state exactly like iterating ``DataLoader(dataset, shuffle=shuffle)`` once:
    the iterator's base seed first (torch/utils/data/dataloader.py `_BaseDataLoaderIter.__init__`), then — when shuffling —
    the sampler's own seed and ``torch.randperm`` on a private generator (torch/utils/data/sampler.py `RandomSampler.__iter__`)."""
    torch.empty((), dtype=torch.int64).random_(generator=generator)  # _base_seed
    if not shuffle:
        return torch.arange(n, dtype=torch.int64)
    if generator is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        gen = torch.Generator()
        gen.manual_seed(seed)
    else:
        gen = generator
    return torch.randperm(n, generator=gen)


class DeviceLoader:
    """Iterable over (coarse, fine) device batches of an HBM-resident :class:`NetCDFSR`; stands in for
    ``torch.utils.data.DataLoader(dataset, batch_size, shuffle)`` (``stage.py:74-81``).  ``len()`` = number of batches."""

    def __init__(self, dataset: NetCDFSR, batch_size: int = 1, shuffle: bool = False, drop_last: bool = False,
                 generator: Optional[torch.Generator] = None) -> None:
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.drop_last = bool(drop_last)
        self.generator = generator
        self.last_permutation: Optional[torch.Tensor] = None  # CPU int64, for tests / reproducibility records

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _resident(self) -> Tuple[torch.Tensor, torch.Tensor]:
        c, f = self.dataset.coarse, self.dataset.fine
        for name, t in (("coarse", c), ("fine", f)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise _lib.DgError(f"DeviceLoader needs the {name} tensor resident on the GPU as contiguous float32 "
                                   f"(got {t.dtype} on {t.device}); the reference moves both there in stage.py:29-32")
        return c, f

    def batch_indices(self) -> List[torch.Tensor]:
        """Draws the epoch's permutation and splits it into batches (CPU tensors)."""
        perm = epoch_permutation(len(self.dataset), self.shuffle, self.generator)
        self.last_permutation = perm
        nb = len(self)
        return [perm[i * self.batch_size:(i + 1) * self.batch_size] for i in range(nb)]

    def skip_first_batch(self) -> None:
        """RNG side effect of ``next(iter(dataloader))`` (the reference does this twice per epoch for its plots,
        wasserstein.py:155,175): keeps later epochs' shuffles aligned with the reference's stream."""
        epoch_permutation(len(self.dataset), self.shuffle, self.generator)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        coarse, fine = self._resident()
        lib = _lib.load()
        batches = self.batch_indices()
        dev = fine.device
        with torch.cuda.device(dev):
            perm_dev = self.last_permutation.to(dev, non_blocking=False)  # one upload per epoch
            n = len(self.dataset)
            ce, fe = coarse[0].numel(), fine[0].numel()
            off = 0
            for idx in batches:
                b = idx.numel()
                cb = torch.empty((b,) + tuple(coarse.shape[1:]), device=dev, dtype=torch.float32)
                fb = torch.empty((b,) + tuple(fine.shape[1:]), device=dev, dtype=torch.float32)
                ip = perm_dev.data_ptr() + 8 * off
                _lib.check(lib.dg_gather_rows(coarse.data_ptr(), ip, b, ce, n, cb.data_ptr(), _lib.stream_ptr()))
                _lib.check(lib.dg_gather_rows(fine.data_ptr(), ip, b, fe, n, fb.data_ptr(), _lib.stream_ptr()))
                off += b
                yield cb, fb
        if self.shuffle and self.generator is not None:
            # a fully consumed RandomSampler draws (and discards) one more permutation for its `num_samples % n` == 0 tail;
            # only visible when the caller shares `generator` with other consumers
            torch.randperm(len(self.dataset), generator=self.generator)
