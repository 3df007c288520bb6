#!/usr/bin/env python
"""Does the operand layout cost tensor throughput?  The 512 <-> C passes of the step as cc_gemm_tc calls (tf32 and bf16),
each timed with its operands K-major and MN-major (the data is laid out accordingly beforehand; the result is the same
matrix).  CUDA events, L2 flushed between launches, median of 10.  One JSON line per case.

    python profiles/gemm_layout_bench.py
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200 import _lib  # noqa: E402
from cubecobrarecommender_b200.ml import tensorcore as TC  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
C, H, B = 20884, 512, 4096
CP = 20992


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for prec, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    mk = lambda *s: (torch.randn(s, device=dev, generator=g) * 0.1).to(dt)
    cases = []
    # forward: z[B][C] = h[B][H] W[H][C]
    h, w_kn, w_nk = mk(B, H), mk(H, CP)[:, :C], mk(C, H)
    z = torch.empty((B, CP), device=dev)[:, :C]
    cases.append(("fwd  M=4096 N=C K=512, A K-major, B MN-major ([K][N], the Keras kernel)", lambda: TC.gemm(h, w_kn, z, precision=prec)))
    cases.append(("fwd  M=4096 N=C K=512, A K-major, B K-major ([N][K], a transposed kernel copy)", lambda: TC.gemm(h, w_nk, z, transb=True, precision=prec)))
    # dX: g[B][H] = dz[B][C] W^T: B = W [H][C] used as [N=H][K=C] (K-major) -- or W^T stored [C][H] used as [K][N] (MN-major)
    dz = mk(B, CP)[:, :C]
    gx = torch.empty((B, H), device=dev)
    cases.append(("dX   M=4096 N=512 K=C, A K-major, B K-major", lambda: TC.gemm(dz, w_kn, gx, transb=True, precision=prec)))
    cases.append(("dX   M=4096 N=512 K=C, A K-major, B MN-major", lambda: TC.gemm(dz, w_nk, gx, precision=prec)))
    # dW: gw[H][C] = h^T[H][B] dz[B][C]: A = h [B][H] used as A^T (MN-major), B = dz [K=B][N=C] (MN-major)
    gw = torch.empty((H, CP), device=dev)[:, :C]
    ht, dzt = mk(H, B), mk(C, B)
    cases.append(("dW   M=512 N=C K=4096, A MN-major, B MN-major (as the step runs it)", lambda: TC.gemm(h, dz, gw, transa=True, precision=prec)))
    cases.append(("dW   M=512 N=C K=4096, A K-major, B K-major (both operands transposed beforehand)", lambda: TC.gemm(ht, dzt, gw, transb=True, precision=prec)))
    for name, fn in cases:
        ms = timed(fn)
        print(json.dumps({"precision": prec, "case": name, "ms": ms, "TFLOPs": 2.0 * B * H * C / ms / 1e9,
                          "mn3": os.environ.get("CC_GEMM_MN3", "1"),
                          "operands_through_3d_maps_so_far": int(_lib.load().cc_gemm_tc_mn3_count())}), flush=True)
