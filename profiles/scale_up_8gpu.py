#!/usr/bin/env python
"""BASELINE configs[4] as written: 1M cubes x 100k cards over the GPUs of one box, cube-sharded (125 000 cubes per rank at
8 ranks), private int32 (C, C) counts by the tensor-core count kernel, then the exchange step in BOTH forms:

* ``allreduce``      ONE all_reduce(SUM) of the 40 GB counts over NVLink (exact), row-normalise to a replicated float32
                     M-hat on every rank (the form BASELINE.json words);
* ``reduce_scatter`` reduce_scatter(SUM) by row block (half the traffic), every rank normalises only its C/8 rows
                     (graph.reduce_scatter_counts + graph.normalise_rows): M-hat comes out row-sharded, which is how
                     the full-I regulariser consumes it, and the column masses are all_reduced (C doubles).

Launch:

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
counts = G.alloc_counts(C, world, dev)                                # (C, C) int32, storage padded to `world` equal row blocks
MODE = os.environ.get("SCALE_UP_MODE", "both")
blk = -(-C // world)
mhat = torch.empty((C, C), dtype=torch.float32, device=dev) if MODE != "reduce_scatter" else None   # outside the timed region
mhat_rows = torch.empty((blk, C), dtype=torch.float32, device=dev)
G.count_cooccurrence(indptr[:4097], indices[:4096 * S], 4096, C, counts=counts, workspace=ws, method="tensor")   # warm-up
G.normalise_rows(counts[:64], 0, C, want_mhat=True, mhat=mhat_rows[:64])     # (first use loads the kernels: ~30 ms)
if mhat is not None:
    G.normalise(counts[:64, :64].contiguous(), want_m64=False, want_mhat=True, want_neg=True)
if world > 1:
    warm = torch.ones(1 << 20, dtype=torch.int32, device=dev)
    dist.all_reduce(warm)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
results = {}
for mode in (("reduce_scatter", "allreduce") if MODE == "both" else (MODE,)):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    G.count_cooccurrence(indptr, indices, K, C, counts=counts, workspace=ws, method="tensor")
    ev[1].record()
    if mode == "allreduce":
        if world > 1:
            dist.all_reduce(counts.view(-1), op=dist.ReduceOp.SUM)
        ev[2].record()
        gr = G.normalise(counts, want_m64=False, want_mhat=True, want_neg=True, mhat=mhat)
        my_diag = counts.diagonal().sum(dtype=torch.int64) if rank == 0 else torch.zeros((), dtype=torch.int64, device=dev)
    else:
        rows, r0 = G.reduce_scatter_counts(counts, rank, world)
        ev[2].record()
        gr = G.normalise_rows(rows, r0, C, want_mhat=True, mhat=mhat_rows[:rows.shape[0]])
        my_diag = rows[:, r0:r0 + rows.shape[0]].diagonal().sum(dtype=torch.int64)
    ev[3].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) / 1e3 for i in range(3)], dtype=torch.float64, device=dev)
    neg_ok = torch.tensor([float(abs(gr.neg_sampler.sum().item() - 1.0) < 1e-9)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(my_diag, op=dist.ReduceOp.SUM)
        dist.all_reduce(neg_ok, op=dist.ReduceOp.MIN)
    # size-independent check of the summed counts: diag[c] = number of cubes holding card c, so the trace equals the
    # number of DISTINCT (cube, card) pairs over all ranks' cubes
    uniq = torch.tensor([0], dtype=torch.int64, device=dev)
    for k0 in range(0, K, 8192):
        k1 = min(k0 + 8192, K)
        seg = indices[k0 * S:k1 * S].view(k1 - k0, S).long().sort(dim=1).values
        uniq += (k1 - k0) + (seg[:, 1:] != seg[:, :-1]).sum()
    if world > 1:
        dist.all_reduce(uniq, op=dist.ReduceOp.SUM)
    t_cnt, t_x, t_norm = t.tolist()
    payload = 4.0 * C * C
    results[mode] = {
        "count_seconds": t_cnt, "exchange_seconds": t_x, "normalise_seconds": t_norm, "build_seconds": t_cnt + t_x + t_norm,
        "cubes_per_s": K_TOTAL / (t_cnt + t_x + t_norm),
        "exchange_busbw_GBps": ((2.0 if mode == "allreduce" else 1.0) * (world - 1) / world * payload / t_x / 1e9)
        if world > 1 and t_x > 0 else None,
        "trace_equals_distinct_pairs": bool(int(my_diag.item()) == int(uniq.item())),
        "neg_sampler_sums_to_one": bool(neg_ok.item()),
        "output": "replicated float32 M-hat (C, C)" if mode == "allreduce" else f"row-sharded float32 M-hat ({blk}, C) per rank"}
t_cnt = results[next(iter(results))]["count_seconds"]
if rank == 0:
    print(json.dumps({
        "modes": results,
        "workload": f"configs[4]: {K_TOTAL} cubes x {C} cards, s={S}, {world} rank(s), {K} cubes per rank",
        "counts_payload_GB": 4.0 * C * C / 1e9,
        "count_dense_equivalent_pops_per_rank": 2.0 * K * C * C / t_cnt / 1e15,
        "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "timing": "CUDA events, max over ranks"}))
if world > 1:
    dist.destroy_process_group()
