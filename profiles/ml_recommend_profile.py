#!/usr/bin/env python
"""BASELINE configs[3] on one GPU, taken apart: 100 000 cubes top-50 through MLRecommender.recommend_device (pinned
host CSR in, ids on the device), chunk sizes 2048 / 4096 / 8192.  Per chunk size one JSON line with the CUDA-event
time of the whole call, the host time spent ENQUEUEING it (the call returns before the GPU finishes: if enqueue time
~ device time the path is launch-bound), and -- for one warm chunk -- CUDA-event times of its stages (first layer,
encoder, decoder small layers, 512 -> C GEMM, masked select).

    python profiles/ml_recommend_profile.py [--cubes 100000] [--trained]
`--trained` spreads the logits like a trained model's (decoder output bias drawn from N(-4, 3)): with the random
initial weights every logit of a row lies within a few hundredths of the others.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200.graph import topn_masked  # noqa: E402
from cubecobrarecommender_b200.ml import inference as INF, model as M  # noqa: E402
from cubecobrarecommender_b200.ml.model import SparseBatch  # noqa: E402
from cubecobrarecommender_b200.workload import make_cubes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cubes", type=int, default=100000)
    ap.add_argument("--trained", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    C, K = 20884, args.cubes
    csr = make_cubes(K, C, cfg=4).pin_memory()
    model = M.CC_Recommender(C, device=dev, seed=0, precision="tf32")
    if args.trained:
        w = model.get_weights_dict()
        rng = np.random.default_rng(0)
        w["main_reconstruction/bias"] = (rng.standard_normal(C) * 3 - 4).astype(np.float32)
        model.set_weights_dict(w)
    for chunk in (2048, 4096, 8192):
        rec = INF.MLRecommender(model, chunk=chunk)
        rec.recommend_device(csr, 50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        ids, vals, cnt = rec.recommend_device(csr, 50)
        e1.record()
        t_enqueue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_dev = e0.elapsed_time(e1) / 1e3
        # one warm chunk, stage by stage
        sub = csr.rows(np.arange(chunk))
        idx = torch.from_numpy(sub.indices).to(dev); ptr_ = torch.from_numpy(sub.indptr).to(dev)
        sb = SparseBatch(idx, ptr_[:-1], (ptr_[1:] - ptr_[:-1]).to(torch.int32))
        out = (ids[:chunk], vals[:chunk], cnt[:chunk])
        stages = {}

        def timed(name, fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                r = fn()
            b.record(); torch.cuda.synchronize()
            stages[name] = a.elapsed_time(b) / reps
            return r
        h = timed("encode (first layer + 3 GEMMs)", lambda: model._encode(sb))
        z = timed("decode (3 small GEMMs + 512->C GEMM)", lambda: model._decode(h, "main"))
        timed("masked top-50 select (fused sigmoid)", lambda: topn_masked(z, ptr_, idx, 50, sigmoid=True, out=out))
        print(json.dumps({"workload": f"ml_recommend top-50, {K} cubes, C={C}, chunk {chunk}, "
                                      f"{'spread (trained-like)' if args.trained else 'random-init'} logits",
                          "device_seconds": t_dev, "recs_per_s": K / t_dev, "host_enqueue_seconds": t_enqueue,
                          "chunks": -(-K // chunk), "one_chunk_stage_ms": {k: round(v, 4) for k, v in stages.items()},
                          "one_chunk_stage_sum_ms": round(sum(stages.values()), 4),
                          "all_counts_full": bool((cnt == 50).all().item())}), flush=True)


if __name__ == "__main__":
    main()
