"""Oracle (TEST INFRASTRUCTURE): co-occurrence graph, M-hat and graph top-N.

CPU restatement of
  * ``create_adjacency_matrix``  -- reference ``src/non_ml/utils.py:75-92``
  * ``y_mtx`` (M-hat)            -- reference ``src/ml/train.py:69-71``
  * ``neg_sampler``              -- reference ``src/ml/generator.py:30``
  * ``simple_recs``              -- reference ``src/scripts/recommend.py:7-18``
  * ``simple_cuts``              -- reference ``src/scripts/cut_cards.py:7-18``

Pinned against the unmodified reference functions by
``tests/golden/make_golden.py`` -> ``tests/golden/graph_small.npz``.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------- counts / M
def cooc_counts(indptr: np.ndarray, indices: np.ndarray, num_cards: int) -> np.ndarray:
    """Exact integer co-occurrence counts ``cnt = X^T X`` (int64, (C, C)).

    Row ``i`` equals ``cubes[cubes[:, i] == 1].sum(0)`` of reference
    ``utils.py:82-84`` (sums of 0/1 values are exact in float64)."""
    import scipy.sparse as sp
    k = len(indptr) - 1
    # (copies: sum_duplicates() compacts the index arrays in place, and the matrix would share the caller's)
    x = sp.csr_matrix((np.ones(len(indices), dtype=np.int64), np.array(indices, copy=True), np.array(indptr, copy=True)),
                      shape=(k, num_cards))
    x.sum_duplicates()
    x.data[:] = 1  # duplicates collapse: build_cubes assigns 1 (utils.py:71)
    return np.ascontiguousarray((x.T @ x).todense(), dtype=np.int64)


def cooc_counts_blocked(indptr: np.ndarray, indices: np.ndarray, num_cards: int, block: int = 3072,
                        out_dtype=np.int32) -> np.ndarray:
    """The same ``cnt = X^T X`` as :func:`cooc_counts` for BASELINE-sized inputs (K = 20 000, C = 21 000), where the
    sparse product needs minutes and ~10 GB: dense float32 column blocks of X multiplied on all host cores.  A count is
    a sum of at most K products of 0/1 values; K < 2**24, so every partial sum is an integer float32 holds exactly and
    the result is exact whatever the summation order.  Only the upper-triangular blocks are computed and mirrored
    (X^T X is symmetric).  ``tests/test_oracle_graph.py`` holds it equal to :func:`cooc_counts`."""
    import torch
    k = len(indptr) - 1
    if k >= 1 << 24:
        raise ValueError("float32 partial sums are exact only below 2**24 cubes")
    x = torch.zeros((k, num_cards), dtype=torch.float32)
    x[torch.from_numpy(np.repeat(np.arange(k), np.diff(indptr))), torch.from_numpy(indices.astype(np.int64))] = 1.0
    cnt = np.zeros((num_cards, num_cards), dtype=out_dtype)      # (duplicates collapse: assignment, utils.py:71)
    starts = list(range(0, num_cards, block))
    for bi, lo in enumerate(starts):
        hi = min(lo + block, num_cards)
        xi_t = x[:, lo:hi].t().contiguous()
        for lo2 in starts[bi:]:
            hi2 = min(lo2 + block, num_cards)
            blk = (xi_t @ x[:, lo2:hi2]).numpy()
            cnt[lo:hi, lo2:hi2] = blk
            if lo2 != lo:
                cnt[lo2:hi2, lo:hi] = blk.T
    return cnt


def adjacency_from_counts(cnt: np.ndarray, force_diag=None) -> np.ndarray:
    """``M[i,:] = cnt[i,:]/cnt[i,i]`` if ``cnt[i,i] != 0`` else ``cnt[i,:]``
    (reference ``utils.py:85-89``); optional ``fill_diagonal`` (``:90-91``)."""
    cnt = np.asarray(cnt)
    diag = np.diagonal(cnt).astype(np.float64)
    m = np.ascontiguousarray(cnt, dtype=np.float64).copy()
    nz = diag != 0
    m[nz] = m[nz] / diag[nz, None]
    if force_diag is not None:
        np.fill_diagonal(m, force_diag)
    return m


def create_adjacency_matrix_loop(cubes: np.ndarray, force_diag=None) -> np.ndarray:
    """Line-by-line restatement of reference ``utils.py:75-92`` (dense float64
    ``cubes``); O(nnz*C) -- small cases and the CPU-baseline timing only."""
    num_cards = cubes.shape[1]
    adj = np.empty((num_cards, num_cards))
    for i in range(num_cards):
        sel = np.where(cubes[:, i] == 1)            # utils.py:82
        step1 = cubes[sel].sum(0)                   # utils.py:83-84
        adj[i] = step1 / step1[i] if step1[i] != 0 else step1   # utils.py:85-89
    if force_diag is not None:
        np.fill_diagonal(adj, force_diag)           # utils.py:90-91
    return adj


def create_adjacency_matrix(cubes: np.ndarray, force_diag=None) -> np.ndarray:
    """Same result as :func:`create_adjacency_matrix_loop`, computed as
    ``X^T X / diag`` (bit-identical in float64: SURVEY.md §4 probe)."""
    from cubecobrarecommender_b200.synth import dense_to_csr
    indptr, indices = dense_to_csr(cubes)
    return adjacency_from_counts(cooc_counts(indptr, indices, cubes.shape[1]), force_diag)


# -------------------------------------------------------------------- M-hat
def m_hat(adj_mtx: np.ndarray) -> np.ndarray:
    """Reference ``train.py:69-71``: copy, diagonal <- 1, each row / its sum."""
    y = adj_mtx.copy()
    np.fill_diagonal(y, 1)
    return y / y.sum(1)[:, None]


def m_hat_rows(indptr: np.ndarray, indices: np.ndarray, num_cards: int, rows: np.ndarray) -> np.ndarray:
    """Rows ``rows`` of M-hat (float64, (R, C)) for BASELINE-sized inputs, without the (C, C) matrices: exactly the
    reference's arithmetic on those rows -- ``cnt[r, :]`` = sum of the cubes containing card r (``utils.py:82-84``),
    divided by ``cnt[r, r]`` when non-zero (``:85-89``), diagonal <- 1 and the row divided by its sum
    (``train.py:69-71``).  ``tests/test_oracle_graph.py`` holds it equal to ``m_hat(...)[rows]``."""
    import scipy.sparse as sp
    k = len(indptr) - 1
    rows = np.asarray(rows, dtype=np.int64)
    x = sp.csr_matrix((np.ones(len(indices)), np.array(indices, copy=True), np.array(indptr, copy=True)),
                      shape=(k, num_cards))
    x.sum_duplicates()
    x.data[:] = 1
    cnt = np.asarray((x.T.tocsr()[rows] @ x).todense(), dtype=np.float64)          # exact integers
    at = np.arange(len(rows))
    diag = cnt[at, rows].copy()
    nz = diag != 0
    cnt[nz] = cnt[nz] / diag[nz, None]
    cnt[at, rows] = 1.0
    return cnt / cnt.sum(1)[:, None]


def neg_sampler(y_mtx: np.ndarray) -> np.ndarray:
    """Reference ``generator.py:30``: column mass of M-hat, normalised."""
    return y_mtx.sum(0) / y_mtx.sum()


# ---------------------------------------------------------------- graph top-N
def simple_recs(cube: np.ndarray, adj_mtx: np.ndarray, int_to_card=None, *, stable=True):
    """Reference ``recommend.py:7-18``.  ``stable=True`` applies the repo's tie
    rule (descending score, ties -> larger index first), i.e.
    ``argsort(kind='stable')[::-1]``; ``stable=False`` is the reference's
    default (unspecified tie order) sort."""
    contains = np.where(cube == 1)[0]
    missing = np.where(cube == 0)[0]
    sub = adj_mtx[contains][:, missing]
    order = sub.sum(0).argsort(kind="stable" if stable else None)[::-1]
    ids = [missing[i] for i in order]
    return ids if int_to_card is None else [int_to_card[i] for i in ids]


def simple_recs_scores(cube: np.ndarray, adj_mtx: np.ndarray) -> np.ndarray:
    """The float64 scores ``simple_recs`` ranks (``recommend.py:10-13``), as a
    full length-C vector (in-cube entries = -inf)."""
    contains = np.where(cube == 1)[0]
    missing = np.where(cube == 0)[0]
    out = np.full(adj_mtx.shape[1], -np.inf)
    out[missing] = adj_mtx[contains][:, missing].sum(0)
    return out


def simple_cuts(cube: np.ndarray, adj_mtx: np.ndarray, int_to_card=None, *, stable=True):
    """Reference ``cut_cards.py:7-18`` (mutates ``adj_mtx``'s diagonal, like
    the reference does at ``:8``).  Tie rule for ``stable=True``: ascending
    score, ties -> smaller index first (``argsort(kind='stable')``)."""
    np.fill_diagonal(adj_mtx, 0)
    contains = np.where(cube == 1)[0]
    sub = adj_mtx[contains][:, contains]
    order = sub.sum(0).argsort(kind="stable" if stable else None)
    ids = [contains[i] for i in order]
    return ids if int_to_card is None else [int_to_card[i] for i in ids]
