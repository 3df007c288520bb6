#!/usr/bin/env python
"""BASELINE configs[4] (1M cubes x 100k cards on 8 GPUs), ONE rank's share on one GPU: 125 000 cubes x 100 000 cards
counted into a private int32 (C, C) matrix (40 GB) by the tensor-core count kernel (4 passes of the 32 768-cube byte
matrix), then row-normalised to float32 M-hat.  Cubes are uniform random card draws generated on the device (the work
of the count GEMM does not depend on the data); correctness at this card count is covered by
tests/test_gpu_graph.py::test_scale_up_card_count_properties.  Prints one JSON line."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200 import _lib, graph as G  # noqa: E402

K, C, S = 125_000, 100_000, 540
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
indices = torch.randint(0, C, (K * S,), dtype=torch.int32, device=dev, generator=g)
indptr = torch.arange(K + 1, dtype=torch.int64, device=dev) * S
lib = _lib.load()
ws = torch.empty(lib.cc_cooc_tc_workspace_bytes(K, C), dtype=torch.uint8, device=dev)
counts = torch.empty((C, C), dtype=torch.int32, device=dev)
mhat = torch.empty((C, C), dtype=torch.float32, device=dev)           # allocated outside the timed region
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
G.count_cooccurrence(indptr[:4097], indices[:4096 * S], 4096, C, counts=counts, workspace=ws, method="tensor")   # warm-up
torch.cuda.synchronize()
ev[0].record()
G.count_cooccurrence(indptr, indices, K, C, counts=counts, workspace=ws, method="tensor")
ev[1].record()
gr = G.normalise(counts, want_m64=False, want_mhat=True, want_neg=True, mhat=mhat)
ev[2].record()
torch.cuda.synchronize()
t_cnt, t_norm = ev[0].elapsed_time(ev[1]) / 1e3, ev[1].elapsed_time(ev[2]) / 1e3
diag_ok = bool((counts.diagonal() > 0).all().item())
sym_i = torch.randint(0, C, (100000,), device=dev); sym_j = torch.randint(0, C, (100000,), device=dev)
sym_ok = bool(torch.equal(counts[sym_i, sym_j], counts[sym_j, sym_i]))
print(json.dumps({"workload": f"one rank of configs[4]: K={K} cubes x C={C} cards, s={S}", "count_seconds": t_cnt,
                  "count_dense_equivalent_pops": 2.0 * K * C * C / t_cnt / 1e15, "normalise_seconds": t_norm,
                  "normalise_GBps": (4.0 + 4.0 + 4.0) * C * C / t_norm / 1e9, "counts_GB": 4.0 * C * C / 1e9,
                  "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "diag_positive": diag_ok,
                  "symmetric_on_sample": sym_ok,
                  "note": "the 8-GPU build adds one all_reduce of the 40 GB int32 counts (exact) before normalise"}))
