"""CPU restatement of the reference's batch assembly.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference being restated (paths under /root/reference/DoWnGAN):
  * GAN/dataloader.py:6-33   NetCDFSR: ``__len__`` = fine.size(0); ``__getitem__(idx)`` -> (coarse[idx], fine[idx])
  * GAN/stage.py:73-81       torch.utils.data.DataLoader(dataset, batch_size=hp.batch_size, shuffle=True) for both the
                             train and the test set (default collate = torch.stack, drop_last False)

The sampling arithmetic lives in torch (torch.utils.data.RandomSampler / BatchSampler, pinned by requirements.txt:6 to
torch 1.12, installed 2.11): the oracle therefore simply RUNS torch's own DataLoader over the restated dataset and records
which rows each batch holds; ``tests/test_dataloader.py`` compares the product loader's index stream with it.
"""
from __future__ import annotations

from typing import List

import torch
from torch.utils.data import DataLoader, Dataset


class NetCDFSR(Dataset):
    """dataloader.py:6-33."""

    def __init__(self, coarse: torch.Tensor, fine: torch.Tensor, device=None):
        self.fine = fine
        self.coarse = coarse

    def __len__(self):
        return self.fine.size(0)

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return self.coarse[idx, ...], self.fine[idx, ...]


def reference_loader(coarse: torch.Tensor, fine: torch.Tensor, batch_size: int, shuffle: bool = True) -> DataLoader:
    """stage.py:73-81."""
    return DataLoader(dataset=NetCDFSR(coarse, fine), batch_size=batch_size, shuffle=shuffle)


def epoch_batches(loader: DataLoader) -> List[tuple]:
    """One epoch of the reference loader: list of (coarse, fine) batches."""
    return [(c, f) for c, f in loader]
