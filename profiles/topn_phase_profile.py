#!/usr/bin/env python
"""Per-phase clock64() breakdown of topn_rowselect_kernel (diagnostic instantiation, cc_topn_rowselect_profile) on the
topn_bench.py workload: 4096 logit rows of C = 20 884, top-50, in-cube cards masked, fused sigmoid.  For each launch
shape prints the mean cycles per cube that thread 0 of a CTA spends in every phase (all CTAs, all cubes)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200 import _lib  # noqa: E402
from cubecobrarecommender_b200._lib import call, ptr, stream_ptr  # noqa: E402

C, LD, N, S = 20884, 20992, 50, 540
PHASES = ["loop_top_prefetch", "wait_row", "sentinels_fence_barrier", "sweep1", "leaders_threshold", "sweep2",
          "survivors_to_keys", "issue_next_final_rank", "write_out"]
dev = torch.device("cuda", 0)
lib = _lib.load()
batch = int(os.environ.get("TOPN_BATCH", 4096))
g = torch.Generator(device=dev).manual_seed(batch)
mp = torch.arange(batch + 1, dtype=torch.int64, device=dev) * S
mi = torch.randint(0, C, (batch * S,), dtype=torch.int32, device=dev, generator=g)
logits = torch.randn((batch, LD), device=dev, generator=g) * 3 - 4
ids = torch.empty((batch, N), dtype=torch.int32, device=dev)
vals = torch.empty((batch, N), dtype=torch.float32, device=dev)
cnt = torch.empty(batch, dtype=torch.int32, device=dev)
for variant in (2, 1, 0):
    grid = int(lib.cc_topn_rowselect_profile_grid(batch, variant))
    prof = torch.zeros((grid, 10), dtype=torch.int64, device=dev)
    for _ in range(2):                      # second launch: warm
        call("cc_topn_rowselect_profile", ptr(logits), LD, C, batch, ptr(mp), ptr(mi), N, variant, ptr(ids), ptr(vals),
             ptr(cnt), ptr(prof), stream_ptr())
    torch.cuda.synchronize()
    p = prof.cpu().double()
    cubes = p[:, 9].sum().item()
    per_cube = (p[:, :9].sum(0) / cubes).tolist()
    print(json.dumps({"variant": {0: "1 CTA/SM, 2 row buffers", 1: "2 CTAs/SM, 1 row buffer", 2: "register form (default), 2 CTAs/SM, 1 row buffer"}[variant], "grid": grid,
                      "batch": batch, "cycles_per_cube_total": sum(per_cube),
                      "cycles_per_cube": {k: round(v, 1) for k, v in zip(PHASES, per_cube)}}), flush=True)
