#!/usr/bin/env python
"""The reg tower's softmax-KL kernel alone at BASELINE configs[1] shapes: R = 4096 logit rows of C = 20 884 (row stride
20 992, as the decoder GEMM writes them), target rows gathered from a (C, C) float32 M-hat by sampled row ids, dlogits
written in place, bias-gradient column sums fused.  Times both persistent kernels (variant 0: 512 threads, targets and
column sums in registers; variant 1: the 1024-thread form) with CUDA events, L2 flushed between launches.
Algorithmic bytes per row: 3 * 4 * C (logits in, target row in, dlogits out).  One JSON line per variant."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200._lib import call, ptr, stream_ptr  # noqa: E402

C, LD, R = 20884, 20992, int(os.environ.get("KL_ROWS", 4096))
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
g = torch.Generator(device=dev).manual_seed(1)
mhat = torch.rand((C, C), device=dev, generator=g)
mhat /= mhat.sum(1, keepdim=True)
rows = torch.randint(0, C, (R,), dtype=torch.int32, device=dev, generator=g)
z0 = torch.randn((R, LD), device=dev, generator=g) * 2
table = torch.zeros(C, dtype=torch.float64, device=dev)
call("cc_kl_target_table", ptr(mhat), C, C, C, ptr(table), stream_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
reps = int(os.environ.get("KL_REPS", 10))
rl = torch.zeros(R, dtype=torch.float64, device=dev)
db = torch.zeros(C, device=dev)
NAMES = {0: "softmax_kl_regs_kernel (auto: 768 threads x 7 slots when the row fits, else 512 x 11)",
         2: "softmax_kl_regs_kernel (512 threads x 11 slots)", 1: "softmax_kl_persistent_kernel (1024 threads)"}
for variant in (0, 2, 1):
    call("cc_softmax_kl_set_variant", variant)
    times = []
    for it in range(reps + 1):
        z = z0.clone()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("cc_softmax_kl_fwd_bwd_ex", ptr(z), LD, ptr(mhat), C, ptr(rows), R, C, LD, 0.1 / R, ptr(z), LD, ptr(rl), 1,
             ptr(db), None, 0, ptr(table), stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    gb = R * 3.0 * 4 * C / 1e9
    print(json.dumps({"kernel": NAMES[variant], "rows": R, "C": C, "ms": ms, "min_ms": min(times),
                      "algorithmic_GB": gb, "GBps": gb / ms * 1e3,
                      "frac_of_hbm_peak": gb / ms * 1e3 / peaks.get("hbm_gbs", 6546.9), "kl_mean": float(rl.sum().item() / R)}))
call("cc_softmax_kl_set_variant", 0)
