#!/usr/bin/env python
"""BASELINE configs[3] on N GPUs: 100 000 cubes, top-50 with in-cube masking, cubes sharded over the ranks (inference has
no exchange step: every rank ranks its own cubes with a replica of the model).  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        profiles/ml_recommend_ngpu.py            (or plain `python profiles/ml_recommend_ngpu.py` for one GPU)

Times: CUDA events around MLRecommender.recommend_device on every rank (pinned host CSR in, results left on the device),
max over ranks; throughput = all cubes / that time.  NOT run in round 1 at N > 1 (GPU budget)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200.dist import shard_range  # noqa: E402
from cubecobrarecommender_b200.ml import inference as INF, model as M  # noqa: E402
from cubecobrarecommender_b200.workload import make_cubes  # noqa: E402

C, K, N = 20884, int(os.environ.get("RECOMMEND_CUBES", 100000)), 50
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
full = make_cubes(K, C, cfg=4)                                              # every rank builds the same cubes
model = M.CC_Recommender(C, device=dev, seed=0, precision="tf32")          # same seed: identical replicas
rec = INF.MLRecommender(model, chunk=8192)
lines = []
# one launch measures every N in {1, 2, 4, ..., world}: for a given N only ranks < N take a shard, the others idle
ns = [n for n in (1, 2, 4, 8, 16) if n <= world]
for n in ns:
    active = rank < n
    if active:
        lo, hi = shard_range(K, rank, n)
        csr = full.rows(np.arange(lo, hi)).pin_memory()
        rec.recommend_device(csr, N)                                       # warm-up (allocator, copy stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.zeros(1, dtype=torch.float64, device=dev)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    if active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids, vals, cnt = rec.recommend_device(csr, N)
        e1.record()
        torch.cuda.synchronize()
        t[0] = e0.elapsed_time(e1) / 1e3
        ok[0] = int((cnt == N).all().item())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    lines.append({"workload": f"ml_recommend top-{N}, {K} cubes, C={C}, in-cube masking, {n} GPU(s), cubes sharded",
                  "seconds": t.item(), "recs_per_s": K / t.item(), "n_gpus": n, "all_counts_full": bool(ok.item()),
                  "timing": "CUDA events around recommend_device (pinned CSR upload included), max over ranks"})
if rank == 0:
    base = lines[0]["recs_per_s"]
    for ln in lines:
        ln["scaling_efficiency_vs_1gpu"] = ln["recs_per_s"] / (base * ln["n_gpus"])
        print(json.dumps(ln))
if world > 1:
    dist.destroy_process_group()
