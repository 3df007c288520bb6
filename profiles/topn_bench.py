#!/usr/bin/env python
"""Masked top-50 select alone (BASELINE configs[3]'s HBM-bound half): `batch` logit rows of C = 20 884 (row stride
20 992, as the decoder GEMM writes them), in-cube cards masked, sigmoid fused.  Times the warp-per-cube streaming select
(algo 1) and the two launch shapes of the CTA-per-cube row select (algo 2, 3) with CUDA events, L2 flushed between launches, and checks that the
two return identical ids.  Algorithmic bytes per cube: 4C (scores) + 4s (mask ids) + 8n + 4 (outputs).
Prints one JSON line per (algo, batch)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200 import _lib, graph as G  # noqa: E402

C, LD, N = 20884, 20992, 50
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
reps = int(os.environ.get("TOPN_REPS", 10))
for batch in [int(x) for x in os.environ.get("TOPN_BATCHES", "2048,4096").split(",")]:
    g = torch.Generator(device=dev).manual_seed(batch)
    S = 540                                       # in-cube cards per cube (uniform draws; duplicates collapse in the mask)
    mp = torch.arange(batch + 1, dtype=torch.int64, device=dev) * S
    mi = torch.randint(0, C, (batch * S,), dtype=torch.int32, device=dev, generator=g)
    logits = torch.randn((batch, LD), device=dev, generator=g) * 3 - 4
    view = logits[:, :C]
    out = {}
    for algo in (1, 2, 3, 4):
        _lib.call("cc_topn_set_algo", algo)
        res = G.topn_masked(view, mp, mi, N, sigmoid=True)           # warm-up
        torch.cuda.synchronize()
        times = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            G.topn_masked(view, mp, mi, N, sigmoid=True, out=res)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        out[algo] = [t.clone() for t in res]
        ms = float(np.median(times))
        nbytes = batch * (4 * C + 8 * N + 4) + 4 * int(mi.numel())
        print(json.dumps({"kernel": {1: "topn_warpselect_kernel<sigmoid>", 2: "topn_rowselect_kernel<sigmoid> 1 CTA/SM, 2 row buffers",
                                     3: "topn_rowselect_kernel<sigmoid> 2 CTAs/SM, 1 row buffer, both sweeps over shared memory",
                                     4: "topn_rowselect_kernel<sigmoid, REGS> 2 CTAs/SM, 1 row buffer (register form, the default: one diverged block per warp, lane-neighbour ranking)"}[algo],
                          "batch": batch, "C": C, "ld": LD, "n": N, "ms": ms, "min_ms": float(min(times)),
                          "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / ms / 1e6,
                          "frac_of_hbm_peak": nbytes / ms / 1e6 / float(peaks.get("hbm_gbs", 6546.9)),
                          "cubes_per_s": batch / ms * 1e3}), flush=True)
    _lib.call("cc_topn_set_algo", 0)
    same = all(all(torch.equal(x, y) for x, y in zip(out[1], out[k])) for k in out if k != 1)
    print(json.dumps({"batch": batch, "identical_ids_vals_counts": same}), flush=True)
