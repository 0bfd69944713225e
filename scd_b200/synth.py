"""Seeded synthetic inputs of the BASELINE.json configs: L2-normalised DINO/GCD-like image
features, CLIP-like bf16-rounded features and vocabulary, random-init centroids.  Pure torch-CPU
generation, so the oracle, the tests and the benchmark all see identical numbers."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

D = 768
TOPK = 5
NUM_COMMON_VOTE = 20
NUM_COMMON_LINEAR = 4


@dataclass(frozen=True)
class Config:
    name: str
    n: int
    k: int
    v: int
    seed: int
    note: str = ''


CONFIGS = {
    'C1': Config('C1', 6_000, 200, 11_000, 1001, 'CUB-200 scale, the reference-CPU-runnable case'),
    'C2': Config('C2', 127_000, 100, 21_000, 1002, 'ImageNet-100 scale, 21k ImageNet-21k vocabulary (headline)'),
    'C3': Config('C3', 20_000, 120, 21_000, 1003, 'Stanford Dogs scale, partially supervised (TE vocabulary size chosen = 21k)'),
    'C4': Config('C4', 127_000, 100, 82_000, 1004, 'WordNet-noun vocabulary sweep (column-sharded at 1/2/4/8)'),
    'C5': Config('C5', 1_280_000, 1000, 100_000, 1005, 'ImageNet-1k scale stress (data+vocab sharded across 8)'),
}


def unit_rows(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True)


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def image_feats(n: int, k_true: int, seed: int, d: int = D, spread: float = 4.0, chunk: int = 65536, y=None):
    """``normalize(randn(N,d) + spread * mu[y])`` with ``mu = normalize(randn(K,d))``, ``y`` uniform - fp32.
    Generated in row chunks so the 1.28M-row config needs no 2x temporary."""
    g = torch.Generator().manual_seed(seed)
    mu = unit_rows(torch.randn(k_true, d, generator=g))
    y_drawn = torch.randint(0, k_true, (n,), generator=g)
    y = y_drawn if y is None else y
    x = torch.empty(n, d, dtype=torch.float32)
    for lo in range(0, n, chunk):
        hi = min(lo + chunk, n)
        x[lo:hi] = unit_rows(torch.randn(hi - lo, d, generator=g) + spread * mu[y[lo:hi]])
    return x, y


def vocabulary(v: int, seed: int, d: int = D, chunk: int = 16384) -> torch.Tensor:
    """CLIP-like zero-shot weights in the reference layout ``[d, V]`` (V contiguous), bf16-rounded fp32."""
    g = torch.Generator().manual_seed(seed)
    w = torch.empty(d, v, dtype=torch.float32)
    for lo in range(0, v, chunk):
        hi = min(lo + chunk, v)
        w[:, lo:hi] = bf16_round(unit_rows(torch.randn(hi - lo, d, generator=g))).t()
    return w


def random_init_centers(x: torch.Tensor, k: int, seed: int) -> torch.Tensor:
    """The reference's 'random' init (``faster_mix_k_means_pytorch.py:45-49``) with ``RandomState(seed)``."""
    idx = np.random.RandomState(seed).choice(len(x), k, replace=False)
    return x[torch.as_tensor(idx)].clone()


def make(cfg: Config, n_rows: int | None = None, v: int | None = None, d: int = D):
    """Inputs of one naming round: dino feats X (fp32), clip feats Xc (bf16-rounded fp32), vocabulary W [d,V],
    initial centroids C0 and the generating classes y."""
    n = cfg.n if n_rows is None else n_rows
    vv = cfg.v if v is None else v
    X, y = image_feats(n, cfg.k, cfg.seed, d)
    Xc, _ = image_feats(n, cfg.k, cfg.seed + 500, d, y=y)          # same classes, different embedding space
    Xc = bf16_round(Xc)
    W = vocabulary(vv, cfg.seed + 900, d)
    C0 = random_init_centers(X, min(cfg.k, n), cfg.seed)
    return dict(X=X, Xc=Xc, W=W, C0=C0, y=y, cfg=cfg)
