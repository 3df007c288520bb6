#!/usr/bin/env python
"""BASELINE configs[4] as written: 1M cubes x 100k cards over the GPUs of one box, cube-sharded (125 000 cubes per rank at
8 ranks), private int32 (C, C) counts by the tensor-core count kernel, ONE all_reduce(SUM) of the 40 GB counts over
NVLink (exact), then row-normalise to float32 M-hat on every rank.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/scale_up_8gpu.py

Cubes are uniform random card draws generated on the device (the work does not depend on the data; correctness at this
card count is covered by tests/test_gpu_graph.py::test_scale_up_card_count_properties and, for the sharded sum, by
tests/test_dist_gloo.py).  Times are CUDA-event times, max over ranks.  NOT run in round 1 (GPU budget); the single-rank
share is profiles/scale_up_shard.py."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200 import _lib, graph as G  # noqa: E402

K_TOTAL, C, S = int(os.environ.get("SCALE_UP_CUBES", 1_000_000)), int(os.environ.get("SCALE_UP_CARDS", 100_000)), 540
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
K = K_TOTAL // world
g = torch.Generator(device=dev).manual_seed(5 + rank)
indices = torch.randint(0, C, (K * S,), dtype=torch.int32, device=dev, generator=g)
indptr = torch.arange(K + 1, dtype=torch.int64, device=dev) * S
lib = _lib.load()
ws = torch.empty(lib.cc_cooc_tc_workspace_bytes(K, C), dtype=torch.uint8, device=dev)
counts = torch.empty((C, C), dtype=torch.int32, device=dev)
mhat = torch.empty((C, C), dtype=torch.float32, device=dev)           # allocated outside the timed region
G.count_cooccurrence(indptr[:4097], indices[:4096 * S], 4096, C, counts=counts, workspace=ws, method="tensor")   # warm-up
if world > 1:
    warm = torch.ones(1 << 20, dtype=torch.int32, device=dev)
    dist.all_reduce(warm)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
G.count_cooccurrence(indptr, indices, K, C, counts=counts, workspace=ws, method="tensor")
ev[1].record()
if world > 1:
    dist.all_reduce(counts.view(-1), op=dist.ReduceOp.SUM)
ev[2].record()
gr = G.normalise(counts, want_m64=False, want_mhat=True, want_neg=True, mhat=mhat)
ev[3].record()
torch.cuda.synchronize()
t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) / 1e3 for i in range(3)], dtype=torch.float64, device=dev)
diag_total = counts.diagonal().sum(dtype=torch.int64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
t_cnt, t_ar, t_norm = t.tolist()
if rank == 0:
    print(json.dumps({
        "workload": f"configs[4]: {K_TOTAL} cubes x {C} cards, s={S}, {world} rank(s), {K} cubes per rank",
        "count_seconds": t_cnt, "allreduce_seconds": t_ar, "normalise_seconds": t_norm,
        "build_seconds": t_cnt + t_ar + t_norm, "cubes_per_s": K_TOTAL / (t_cnt + t_ar + t_norm),
        "allreduce_payload_GB": 4.0 * C * C / 1e9,
        "allreduce_busbw_GBps": (2.0 * (world - 1) / world * 4.0 * C * C / t_ar / 1e9) if world > 1 and t_ar > 0 else None,
        "count_dense_equivalent_pops_per_rank": 2.0 * K * C * C / t_cnt / 1e15,
        "diag_sum_equals_total_draw_upper_bound": bool(int(diag_total.item()) <= K_TOTAL * S),
        "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "timing": "CUDA events, max over ranks"}))
if world > 1:
    dist.destroy_process_group()
