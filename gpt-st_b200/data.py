"""Batch assembly without the per-sample collate.

The reference builds its loaders as ``DataLoader(TensorDataset(X, Y), batch_size, shuffle, drop_last)`` on tensors that already
live on the GPU (lib/dataloader.py:92-99).  The default collate then indexes the dataset once per SAMPLE and stacks: 2 x B tiny
slice kernels + 2 ``stack`` kernels per batch (1-2 ms of pure launch overhead at batch 64 -- about a third of one training step
of this implementation).  ``BatchedTensorLoader`` yields the same batches, in the same order for the same seed, with one
``index_select`` per tensor: it consumes the global CPU generator exactly like ``DataLoader`` + ``RandomSampler`` do (one draw
for the iterator's base seed, one for the sampler's seed, then ``randperm`` on a private generator).

    train_loader = BatchedTensorLoader(X, Y, batch_size=64, shuffle=True, drop_last=True)     # instead of data_loader(...)
"""
from __future__ import annotations

import torch


class BatchedTensorLoader:
    def __init__(self, *tensors: torch.Tensor, batch_size: int, shuffle: bool = True, drop_last: bool = True):
        if not tensors or any(t.shape[0] != tensors[0].shape[0] for t in tensors):
            raise ValueError("BatchedTensorLoader: tensors must share their first dimension")
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        self.tensors, self.batch_size, self.shuffle, self.drop_last = tensors, int(batch_size), shuffle, drop_last

    def __len__(self) -> int:
        n = self.tensors[0].shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.tensors[0].shape[0]
        dev = self.tensors[0].device
        # DataLoader draws the iterator's base seed first (even with num_workers == 0) ...
        torch.empty((), dtype=torch.int64).random_()
        if self.shuffle:
            # ... then RandomSampler seeds a private generator from the global one and permutes with it
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            gen = torch.Generator()
            gen.manual_seed(seed)
            order = torch.randperm(n, generator=gen)
        else:
            order = torch.arange(n)
        order = order.to(dev, non_blocking=True)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            yield tuple(t.index_select(0, idx) for t in self.tensors)
