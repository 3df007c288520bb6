"""Seeded synthetic cubes.

Nothing real ships with the reference (data.zip and every model are git-LFS
pointers), so every test and benchmark runs on synthetic cubes with a fixed law:

* card popularity ``p_c ∝ (c+1)^-0.8`` (Zipf-like, heavy tailed),
* cube size ``s_k ~ U{360..720}``,
* a cube = ``s_k`` *distinct* cards drawn without replacement from ``p``
  (exponential race / Gumbel top-k, which is exactly successive sampling
  without replacement),
* seed ``numpy.random.default_rng(20884 + cfg)``.

Cubes are produced directly in CSR form (``indptr`` int64, ``indices`` int32,
sorted inside each cube) -- the dense float64 ``(K, C)`` matrix the reference's
``build_cubes`` makes (reference ``src/non_ml/utils.py:57-73``) is only
materialised on request, for the oracle.
"""
from __future__ import annotations

import numpy as np

ZIPF_EXPONENT = 0.8
SIZE_LO, SIZE_HI = 360, 720
SEED_BASE = 20884


def popularity(num_cards: int, exponent: float = ZIPF_EXPONENT) -> np.ndarray:
    p = (np.arange(num_cards, dtype=np.float64) + 1.0) ** (-exponent)
    return p / p.sum()


def synth_cubes_csr(num_cubes: int, num_cards: int, cfg: int = 1, *,
                    size_lo: int = SIZE_LO, size_hi: int = SIZE_HI,
                    exponent: float = ZIPF_EXPONENT, chunk: int = 512,
                    seed: int | None = None):
    """Return ``(indptr int64 [K+1], indices int32 [nnz])`` of K synthetic cubes."""
    rng = np.random.default_rng(SEED_BASE + cfg if seed is None else seed)
    size_hi = min(size_hi, num_cards)
    size_lo = min(size_lo, size_hi)
    sizes = rng.integers(size_lo, size_hi + 1, size=num_cubes)
    inv_p = (1.0 / popularity(num_cards, exponent)).astype(np.float32)
    indptr = np.zeros(num_cubes + 1, dtype=np.int64)
    np.cumsum(sizes, out=indptr[1:])
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    for lo in range(0, num_cubes, chunk):
        hi = min(lo + chunk, num_cubes)
        # exponential race: the s smallest of E_c / p_c are a draw of s cards
        # without replacement with probabilities proportional to p
        keys = rng.standard_exponential(size=(hi - lo, num_cards), dtype=np.float32)
        keys *= inv_p
        smax = int(sizes[lo:hi].max())
        if smax < num_cards:
            part = np.argpartition(keys, smax - 1, axis=1)[:, :smax]
        else:
            part = np.broadcast_to(np.arange(num_cards), (hi - lo, num_cards)).copy()
        pk = np.take_along_axis(keys, part, axis=1)
        order = np.argsort(pk, axis=1, kind="stable")
        ranked = np.take_along_axis(part, order, axis=1)
        for r in range(hi - lo):
            s = int(sizes[lo + r])
            row = np.sort(ranked[r, :s])
            indices[indptr[lo + r]:indptr[lo + r + 1]] = row
    return indptr, indices


def csr_to_dense(indptr: np.ndarray, indices: np.ndarray, num_cards: int,
                 dtype=np.float64) -> np.ndarray:
    """Dense 0/1 ``(K, C)`` matrix, the layout ``build_cubes`` returns
    (reference ``src/non_ml/utils.py:58,71``)."""
    k = len(indptr) - 1
    x = np.zeros((k, num_cards), dtype=dtype)
    rows = np.repeat(np.arange(k), np.diff(indptr))
    x[rows, indices] = 1
    return x


def dense_to_csr(cubes: np.ndarray):
    """CSR of the entries equal to 1 (the reference tests ``== 1`` everywhere:
    ``utils.py:82``, ``generator.py:83``, ``recommend.py:8``)."""
    cubes = np.asarray(cubes)
    rows, cols = np.nonzero(cubes == 1)
    k = cubes.shape[0]
    counts = np.bincount(rows, minlength=k)
    indptr = np.zeros(k + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, cols.astype(np.int32)


def glorot_uniform(fan_in: int, fan_out: int, gen) -> "np.ndarray":
    """Keras ``glorot_uniform`` kernel of shape (in, out); ``gen`` is a
    ``torch.Generator`` or a numpy Generator."""
    limit = float(np.sqrt(6.0 / (fan_in + fan_out)))
    try:
        import torch
        if isinstance(gen, torch.Generator):
            w = (torch.rand(fan_in, fan_out, generator=gen, dtype=torch.float32) * 2 - 1) * limit
            return w.numpy()
    except ImportError:  # pragma: no cover
        pass
    return ((gen.random((fan_in, fan_out), dtype=np.float32) * 2 - 1) * limit).astype(np.float32)
