"""Batched ML recommendation: D1(E(cube)) -> masked top-N (BASELINE.json configs[3]).

Mirrors the ranking walk of reference ``src/scripts/ml_recommend.py:78-116`` /
``web/ml_recommend_web.py:39-64`` for many cubes at once with the model resident in HBM (the
reference reloads a 390 MB SavedModel per request, web/ml_recommend_web.py:37):

    results   = decoder(encoder(cube))            float32 sigmoid probabilities (C,)
    additions = first `amount` of argsort(results)[::-1] with cube[rec] != 1
    cuts      = results[idx] for every in-cube idx

Ranking is done on the float32 probabilities (not the logits) so that saturated scores tie exactly as
they do in the reference; ties are ordered larger-index-first (``argsort(kind='stable')[::-1]``).
"""
from __future__ import annotations

import numpy as np
import torch

from .._lib import call, ptr, stream_ptr
from ..graph import topn_masked
from ..sparse import CubeCSR
from .model import CC_Recommender, SparseBatch


class MLRecommender:
    def __init__(self, model: CC_Recommender, chunk: int = 2048):
        self.model = model
        self.chunk = int(chunk)

    def probabilities(self, csr: CubeCSR) -> torch.Tensor:
        """(batch, C) float32 sigmoid outputs of ``decoder(encoder(x))``."""
        m = self.model
        sb = SparseBatch.from_csr(csr, m.device)
        z = m._decode(m._encode(sb), "main")
        # sigmoid in place over the whole (padded) buffer the logits view lives in
        full = z._base if z._base is not None else z
        call("cc_sigmoid_f32", ptr(full), ptr(full), full.numel(), stream_ptr())
        return z

    def recommend(self, csr: CubeCSR, amount: int):
        """Returns (add_ids int32 (K, n), add_scores float32 (K, n), counts int32 (K,)) on the host."""
        k = csr.num_cubes
        n = max(1, min(int(amount), csr.num_cards))
        ids = np.full((k, n), -1, dtype=np.int32)
        vals = np.zeros((k, n), dtype=np.float32)
        cnts = np.zeros(k, dtype=np.int32)
        dev = self.model.device
        for lo in range(0, k, self.chunk):
            hi = min(lo + self.chunk, k)
            sub = csr.rows(np.arange(lo, hi))
            probs = self.probabilities(sub)
            mp = torch.from_numpy(sub.indptr).to(dev)
            mi = torch.from_numpy(sub.indices if len(sub.indices) else np.zeros(1, np.int32)).to(dev)
            i_, v_, c_ = topn_masked(probs, mp, mi, n, only_listed=False, descending=True)
            ids[lo:hi] = i_.cpu().numpy(); vals[lo:hi] = v_.cpu().numpy(); cnts[lo:hi] = c_.cpu().numpy()
        return ids, vals, cnts

    def recommend_one(self, cube_indices, amount, int_to_card):
        """The ``{"additions": {...}, "cuts": {...}}`` dict of reference web/ml_recommend_web.py:48-67."""
        num_cards = self.model.N
        csr = CubeCSR.from_lists([cube_indices], num_cards)
        probs = self.probabilities(csr)
        dev = self.model.device
        mp = torch.from_numpy(csr.indptr).to(dev)
        mi = torch.from_numpy(csr.indices if len(csr.indices) else np.zeros(1, np.int32)).to(dev)
        n = max(1, min(int(amount), num_cards))
        ids, vals, cnt = topn_masked(probs, mp, mi, n, only_listed=False, descending=True)
        c = int(cnt[0])
        ids = ids[0, :c].cpu().numpy(); vals = vals[0, :c].cpu().numpy()
        results = probs[0].cpu().numpy()
        output = {"additions": dict(), "cuts": dict()}
        for rec, score in zip(ids, vals):
            output["additions"][int_to_card[int(rec)]] = float(score)
        for idx in cube_indices:                       # cubelist order, duplicates overwrite (dict)
            output["cuts"][int_to_card[idx]] = float(results[idx])
        return output
