"""On-device batch source for the training loop (input pipeline).

The reference's `nMNIST` dataset object (utils/mnist_utils/mnist_addition_dataset.py:16-147) keeps the images as a
numpy array on the host and builds every batch in Python — one `get_image` call per sample, standardisation per
image, `np.array` of a list — after which the training loop copies it to the device
(utils/mnist_utils/train.py:105-120).  `DevicePairsLoader` keeps the whole image tensor in HBM (uint8 or float:
42 000 pairs are 66 MB as uint8) and serves a batch as ONE gather + affine standardisation on the device: no host
array, no H2D copy, no synchronisation per step.  It offers what the loop uses of the reference object: `len()`,
`reset_counter()`, iteration / `__getitem__`, `batch_size`, `worlds`, `world_counter`, `mean`, `std`, and the same
batch composition — every batch holds samples of ONE world (digit pair), worlds are drawn at random without
replacement among those that still have batches left, a world's batches are served last-to-first
(`mnist_addition_dataset.py:89-121`).  Labels are `[d1, ..., dn, sum, -1]` rows (`get_image`, :67-81).

`SyntheticPairsLoader` builds such a source from nothing (there is no network for MNIST here): random images of the
reference's shape, `samples_per_world` per digit pair — what bench.py and the tests use.
"""
from __future__ import annotations

from itertools import product
from typing import Dict, List, Optional, Sequence, Tuple

import torch


class DevicePairsLoader:
    def __init__(self, images: torch.Tensor, labels: torch.Tensor, batch_size: int = 32, digits: int = 10,
                 sequence_len: int = 2, worlds: Optional[Sequence[Tuple[int, ...]]] = None, device=None,
                 idxs: Optional[Dict[Tuple[int, ...], torch.Tensor]] = None, seed: Optional[int] = None,
                 mean: Optional[float] = None, std: Optional[float] = None):
        """images [N, 28, 28*sequence_len] (uint8 or float), labels [N, sequence_len(+1)] digit columns (a trailing sum
        column is ignored and recomputed); `idxs`: per world the sample indices to use (the reference's reproducibility
        hook) — by default every sample of the world, truncated to a multiple of batch_size."""
        device = torch.device(device) if device is not None else images.device
        self.images = images.to(device)
        digits_cols = labels[:, :sequence_len].to(torch.long)
        self.labels = torch.cat([digits_cols, digits_cols.sum(1, keepdim=True),
                                 -torch.ones(len(digits_cols), 1, dtype=torch.long)], 1).to(device)
        self.device, self.batch_size, self.n_digits = device, int(batch_size), digits
        self.worlds: List[Tuple[int, ...]] = list(worlds) if worlds else list(product(range(digits), repeat=sequence_len))
        imf = self.images.float()
        self.mean = float(imf.mean()) if mean is None else float(mean)      # reference: whole-array statistics (:37)
        self.std = float(imf.std(unbiased=False)) if std is None else float(std)
        self._gen = torch.Generator().manual_seed(seed) if seed is not None else None
        # per world: [n_batches, batch_size] sample indices, resident on the device
        self.idxs: Dict[Tuple[int, ...], torch.Tensor] = {}
        host_digits = digits_cols.cpu()
        for c in self.worlds:
            if idxs is not None:
                ids = torch.as_tensor(idxs[c]).reshape(-1)
            else:
                mask = torch.ones(len(host_digits), dtype=torch.bool)
                for pos, d in enumerate(c):
                    mask &= host_digits[:, pos] == d
                ids = torch.nonzero(mask).reshape(-1)
            n = (len(ids) // self.batch_size) * self.batch_size
            self.idxs[c] = ids[:n].reshape(-1, self.batch_size).to(device)
        self.samples_x_world = min((v.numel() for v in self.idxs.values()), default=0)
        for c in self.worlds:   # the reference assumes the same count for every world (:62-64)
            self.idxs[c] = self.idxs[c][: self.samples_x_world // self.batch_size]
        self.n_samples = len(self.worlds) * self.samples_x_world
        self.reset_counter()

    def __len__(self) -> int:
        """number of batches of an epoch (mnist_addition_dataset.py:83-87)"""
        return (len(self.worlds) * self.samples_x_world + self.batch_size - 1) // self.batch_size

    def reset_counter(self) -> None:
        self.world_counter = {c: self.samples_x_world // self.batch_size for c in self.worlds}

    def __getitem__(self, _):
        """-> (images [bs, H, W] float32 standardised, labels [bs, n+2] long), both on the device."""
        remaining = [c for c in self.worlds if self.world_counter[c] > 0]
        if not remaining:
            raise IndexError("DevicePairsLoader exhausted: call reset_counter()")   # the reference raises a str here
        c = remaining[int(torch.randint(len(remaining), (1,), generator=self._gen))]
        self.world_counter[c] -= 1
        ids = self.idxs[c][self.world_counter[c]]
        images = self.images.index_select(0, ids).float().sub_(self.mean).div_(self.std)
        return images, self.labels.index_select(0, ids)

    def __iter__(self):
        for _ in range(len(self)):
            yield self[0]


def SyntheticPairsLoader(samples_per_world: int = 64, batch_size: int = 32, digits: int = 10, sequence_len: int = 2,
                         device="cuda", seed: int = 1234) -> DevicePairsLoader:
    """Random uint8 'images' of the reference's 28 x (28 * sequence_len) shape, `samples_per_world` per digit tuple."""
    g = torch.Generator().manual_seed(seed)
    worlds = list(product(range(digits), repeat=sequence_len))
    labels = torch.tensor(worlds, dtype=torch.long).repeat_interleave(samples_per_world, dim=0)
    images = torch.randint(0, 256, (len(labels), 28, 28 * sequence_len), dtype=torch.uint8, generator=g)
    return DevicePairsLoader(images, labels, batch_size=batch_size, digits=digits, sequence_len=sequence_len,
                             device=device, seed=seed)
