"""MNIST access without a network.

``load_mnist`` reads the standard IDX files through torchvision when they are
on disk (``data_dir/MNIST/raw``) and otherwise returns ``synthetic_mnist`` — a
deterministic 10-class dataset of the same shape/dtype (uint8 28x28, 60k/10k)
whose classes are separable by a small conv net, so accuracy-vs-round curves
remain meaningful on a box with no dataset (there is no egress in the build
environment; reference loader: experiments/dist_mnist_ex.py:98-105).
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

from .shards import Shard

MNIST_MEAN, MNIST_STD = 0.1307, 0.3081


def synthetic_mnist(num: int, seed: int = 0, noise: float = 0.35, classes=None) -> Shard:
    """Class-conditional strokes + noise, uint8 ``[num,1,28,28]`` / int64 labels.
    ``classes`` restricts the labels (e.g. one class per node for the hetero split)."""
    g = torch.Generator().manual_seed(1234)  # prototypes are shared by train and val
    protos = torch.zeros(10, 28, 28)
    for c in range(10):
        for _ in range(4):  # four random strokes per class
            x0, y0 = torch.randint(4, 24, (2,), generator=g).tolist()
            dx, dy = (torch.rand(2, generator=g) * 2 - 1).tolist()
            for t in range(14):
                xi, yi = int(x0 + dx * t), int(y0 + dy * t)
                if 1 <= xi < 27 and 1 <= yi < 27:
                    protos[c, yi - 1: yi + 2, xi - 1: xi + 2] += 0.5
    protos.clamp_(0, 1)
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, 10, num) if classes is None else np.asarray(classes)[rng.integers(0, len(classes), num)]
    combo = rng.integers(0, 25, num)  # one of 25 translations in [-2, 2]^2
    pn = protos.numpy()
    rolled = np.stack([np.roll(pn, (c // 5 - 2, c % 5 - 2), (1, 2)) for c in range(25)])  # [25,10,28,28]
    out = np.empty((num, 28, 28), dtype=np.uint8)
    chunk = 2048  # small blocks: fresh pages are very expensive in sandboxed containers
    for a in range(0, num, chunk):
        sl = slice(a, min(num, a + chunk))
        blk = rolled[combo[sl], labels[sl]]
        blk += noise * rng.random(blk.shape, dtype=np.float32)
        np.clip(blk, 0.0, 1.0, out=blk)
        blk *= 255.0
        out[sl] = blk.astype(np.uint8)
    x = torch.from_numpy(out).unsqueeze(1)
    labels = torch.from_numpy(labels.astype(np.int64))
    return Shard(x, labels, (MNIST_MEAN, MNIST_STD))


def synthetic_mnist_hard(num: int, seed: int = 0, label_noise: float = 0.02, dropout: float = 0.25, shift: int = 3,
                         noise: float = 0.55) -> Shard:
    """A NON-separable 10-class stand-in for MNIST (on ``synthetic_mnist`` every algorithm reaches
    100 %, so accuracy-vs-rounds cannot tell DSGD from DiNNO).  Classes are built from a shared pool of strokes — every
    class shares three of its five strokes with other classes — each sample drops strokes at random, is translated by up
    to +-3 pixels, scaled in intensity, buried in heavier noise, and ``label_noise`` of the labels are wrong.  A
    centralised MNISTConvNet(3,5,64) saturates around 90-95 % here, leaving headroom that separates the algorithms."""
    g = torch.Generator().manual_seed(4321)
    n_pool = 20
    pool = torch.zeros(n_pool, 28, 28)
    for k in range(n_pool):
        x0, y0 = torch.randint(5, 23, (2,), generator=g).tolist()
        ang = float(torch.rand(1, generator=g)) * 6.2832
        dx, dy = float(np.cos(ang)), float(np.sin(ang))
        for t in range(-7, 8):
            xi, yi = int(round(x0 + dx * t)), int(round(y0 + dy * t))
            if 1 <= xi < 27 and 1 <= yi < 27:
                pool[k, yi - 1: yi + 2, xi - 1: xi + 2] += 0.45
    pool.clamp_(0, 1)
    # class c: two private strokes (2c, 2c+1) + three shared with its neighbors in class space
    members = np.asarray([[2 * c, 2 * c + 1, (2 * c + 2) % n_pool, (2 * c + 5) % n_pool, (2 * c + 9) % n_pool] for c in range(10)])
    pn = pool.numpy()
    rng = np.random.default_rng(seed + 7919)
    labels = rng.integers(0, 10, num)
    out = np.empty((num, 28, 28), dtype=np.uint8)
    chunk = 1024
    for a in range(0, num, chunk):
        sl = slice(a, min(num, a + chunk))
        n = sl.stop - sl.start
        keep = rng.random((n, 5)) > dropout                            # stroke dropout
        keep[np.arange(n), rng.integers(0, 5, n)] = True               # at least one stroke survives
        w = keep * rng.uniform(0.6, 1.0, (n, 5))
        blk = np.einsum("nk,nkhw->nhw", w.astype(np.float32), pn[members[labels[sl]]])
        sh = rng.integers(-shift, shift + 1, (n, 2))
        for i in range(n):
            blk[i] = np.roll(blk[i], (sh[i, 0], sh[i, 1]), (0, 1))
        blk += noise * rng.random(blk.shape, dtype=np.float32)
        np.clip(blk, 0.0, 1.0, out=blk)
        out[sl] = (blk * 255.0).astype(np.uint8)
    flip = rng.random(num) < label_noise
    labels = np.where(flip, rng.integers(0, 10, num), labels)
    return Shard(torch.from_numpy(out).unsqueeze(1), torch.from_numpy(labels.astype(np.int64)), (MNIST_MEAN, MNIST_STD))


def load_mnist(data_dir: str, train: bool, synthetic_size: int | None = None,
               allow_synthetic: bool = True, source: str = "auto") -> Tuple[Shard, str]:
    """Returns ``(shard, source)`` with ``source in {"mnist", "synthetic", "synthetic_hard"}``.
    ``source``: ``auto`` (MNIST when on disk, else the separable synthetic set), ``mnist`` (fail when the IDX files are
    missing — what the *_PAPER configs should use on a box that has the data), ``synthetic`` / ``synthetic_hard``."""
    n = synthetic_size or (60000 if train else 10000)
    if source == "synthetic_hard":
        return synthetic_mnist_hard(n, seed=0 if train else 1), "synthetic_hard"
    if source == "synthetic":
        return synthetic_mnist(n, seed=0 if train else 1), "synthetic"
    if source == "mnist":
        allow_synthetic = False
    raw = os.path.join(data_dir or "", "MNIST", "raw")
    if os.path.isdir(raw):
        try:
            from torchvision import datasets

            ds = datasets.MNIST(data_dir, train=train, download=False)
            return Shard(ds.data.unsqueeze(1), ds.targets, (MNIST_MEAN, MNIST_STD)), "mnist"
        except Exception:  # pragma: no cover - corrupt files fall through
            pass
    if not allow_synthetic:
        raise FileNotFoundError(f"MNIST not found under {data_dir!r} and synthetic data disabled")
    n = synthetic_size or (60000 if train else 10000)
    return synthetic_mnist(n, seed=0 if train else 1), "synthetic"
