#!/usr/bin/env python
"""Profile driver: the bench.py train-step workload (configs[1]) with the CUDA profiler range
open around ONE warm step, for `ncu --profile-from-start off`.  Optionally also one graph build
(configs[0] sizes) and one batched top-50 call (configs[3] sizes) inside the range.

    python profiles/step_driver.py [--graph] [--recommend] [--precision tf32]
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from cubecobrarecommender_b200 import _lib, graph as G  # noqa: E402
from cubecobrarecommender_b200.ml import engine as E, inference as INF, model as M  # noqa: E402
from cubecobrarecommender_b200.workload import TRAIN_STEP as W, make_cubes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--recommend", action="store_true")
    ap.add_argument("--warm", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    C, B, R = W["num_cards"], W["batch"], W["reg_rows"]
    csr = make_cubes(W["num_cubes"], C, cfg=W["cfg"] * 1000)
    gr = G.build_graph(csr, dev, want_m64=False, want_mhat=True, want_neg=True)
    prob, alias = E.alias_table(gr.neg_sampler.cpu().numpy(), dev)
    model = M.CC_Recommender(C, device=dev, seed=0, precision=args.precision)
    eng = E.DAEEngine(model, gr.mhat, batch=B, reg_rows=R, reg=W["reg"], max_cube_size=720)
    del gr.counts
    indptr, indices = G.upload_csr(csr, dev)
    ids = torch.arange(B, dtype=torch.int32, device=dev)

    def step():
        eng.sample_batch(indptr, indices, ids, prob, alias, W["noise"], W["noise_std"], seed=1234)
        return eng.train_step()

    for _ in range(args.warm):
        step()
    torch.cuda.synchronize()
    graph_in = rec = None
    if args.graph:
        K2, C2 = 20000, 21000
        csr2 = make_cubes(K2, C2, cfg=1)
        ip2, ix2 = G.upload_csr(csr2, dev)
        lib = _lib.load()
        bits = torch.empty((lib.cc_bits_words(K2), lib.cc_bits_cpad(C2)), dtype=torch.int32, device=dev)
        counts = torch.empty((C2, C2), dtype=torch.int32, device=dev)
        graph_in = (ip2, ix2, K2, C2, counts, bits)
        G.count_cooccurrence(ip2, ix2, K2, C2, counts=counts, bits=bits)
    if args.recommend:
        rec = INF.MLRecommender(model, chunk=2048)
        csr3 = make_cubes(2048, C, cfg=4)
        rec.recommend(csr3, 50)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = step()
    if graph_in is not None:
        ip2, ix2, K2, C2, counts, bits = graph_in
        G.count_cooccurrence(ip2, ix2, K2, C2, counts=counts, bits=bits)
        g2 = G.normalise(counts)
        del g2
    if rec is not None:
        rec.recommend(csr3, 50)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", [float(v) for v in loss.cpu().numpy()])


if __name__ == "__main__":
    main()
