"""N-GPU train step == 1-GPU train step, on a FIXED global batch (SURVEY.md §4 "same step on 1 vs 2/4/8 GPUs").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        -m cubecobrarecommender_b200.dp_check [--precision tf32] [--steps 3] [--modes p2p_unicast,p2p_multicast,nccl]

Every rank first trains ``steps`` steps on the WHOLE global batch by itself (engine with ``data_parallel=False``: the
1-GPU reference), then the ranks train the same steps data-parallel on their B/N cubes and R/N regulariser rows of the
same (x, y, r) -- once per exchange mode -- and rank 0 prints one JSON line with, per mode:

* ``loss_rel_err``       per step, |loss_N - loss_1| / loss_1 of the total loss (bar: 1e-6 on the first step, where both
                         runs hold identical weights; afterwards two slightly different nets are compared);
* ``weights_max_abs_diff`` after the last step (bar: Adam tolerance, a few lr);
* ``replicas_bit_identical``  the parameters of all ranks are equal bit for bit;
* ``adam_state_max_abs_diff`` m and v after ``gather_adam_state()`` against the 1-GPU run (the p2p mode keeps them sliced).

Used by tests/test_gpu_multi.py and by ``bench.py --check``.  There is no oracle in here: the 1-GPU step itself is held
to the oracle by tests/test_gpu_baseline_shapes.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

MODES = {"p2p_unicast": dict(CC_DP_MODE="p2p", CC_P2P_MULTICAST="0"),
         "p2p_multicast": dict(CC_DP_MODE="p2p", CC_P2P_MULTICAST="1"),
         "p2p_overlap": dict(CC_DP_MODE="p2p_overlap"),
         "nccl": dict(CC_DP_MODE="nccl"),
         "nccl_overlap": dict(CC_DP_MODE="nccl_overlap")}


def _bits_equal_across_ranks(t, group=None):
    import torch.distributed as dist
    v = t.view(torch.int32)
    hi, lo = v.clone(), v.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    return bool(torch.equal(hi, lo))


def run_check(precision="tf32", steps=3, modes=("p2p_unicast", "p2p_multicast", "nccl"), num_cards=None,
              global_batch=None, log=lambda m: None):
    import torch.distributed as dist
    from . import graph as G
    from .ml import engine as E, model as M
    from .workload import TRAIN_STEP, make_cubes
    W = TRAIN_STEP
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    c = int(num_cards or W["num_cards"])
    gb = int(global_batch or W["batch"])
    gr_rows = gb
    if gb % world:
        raise ValueError(f"global batch {gb} not divisible by {world} ranks")
    lb = gb // world
    csr = make_cubes(gb, c, cfg=W["cfg"] * 1000)                    # the same cubes on every rank
    graph = G.build_graph(csr, dev, allreduce=False, want_m64=False, want_mhat=True, want_neg=True)
    prob, alias = E.alias_table(graph.neg_sampler.cpu().numpy(), dev)
    indptr, indices = G.upload_csr(csr, dev)

    def fresh_model():
        return M.CC_Recommender(c, device=dev, seed=0, precision=precision)

    # ---- 1-GPU reference on the whole global batch (and the fixed noise output every run uses) ----
    ref_model = fresh_model()
    ref = E.DAEEngine(ref_model, graph.mhat, batch=gb, reg_rows=gr_rows, reg=W["reg"], max_cube_size=720,
                      data_parallel=False)
    ref.sample_batch(indptr, indices, torch.arange(gb, dtype=torch.int32, device=dev), prob, alias, W["noise"],
                     W["noise_std"], seed=4321)
    ref.check_overflow()
    fixed = [ref.x_idx, ref.x_len, ref.y_bits, ref.reg_rows]
    for t in fixed:                                                  # rank 0's draw is THE batch
        dist.broadcast(t, src=0)
    x_idx, x_len, y_bits, reg_rows = [t.clone() for t in fixed]
    ref.set_batch(M.SparseBatch(ref.x_idx.view(-1), ref.x_start, ref.x_len), y_bits, reg_rows[:gr_rows])
    ref_losses = []
    for _ in range(steps):
        ref_losses.append(ref.train_step().cpu().numpy().copy())
    ref_params = ref_model.store.params.clone()
    ref_m, ref_v = ref_model.store.adam_m.clone(), ref_model.store.adam_v.clone()
    del ref, ref_model
    torch.cuda.empty_cache()
    out = {"world": world, "precision": precision, "steps": steps, "num_cards": c, "global_batch": gb,
           "per_rank_batch": lb, "reference_loss": [float(l[2]) for l in ref_losses], "modes": {}}
    saved_env = {k: os.environ.get(k) for k in ("CC_DP_MODE", "CC_P2P_MULTICAST")}
    try:
        for mode in modes:
            for k in saved_env:
                os.environ.pop(k, None)
            os.environ.update(MODES[mode])
            model = fresh_model()
            eng = E.DAEEngine(model, graph.mhat, batch=lb, reg_rows=lb, reg=W["reg"], max_cube_size=720,
                              global_batch=gb, global_reg_rows=gr_rows)
            sl = slice(rank * lb, (rank + 1) * lb)
            eng.x_idx.copy_(x_idx[sl]); eng.x_len.copy_(x_len[sl])
            eng.set_batch(M.SparseBatch(eng.x_idx.view(-1), eng.x_start, eng.x_len), y_bits[sl], reg_rows[sl])
            losses = []
            try:
                for _ in range(steps):
                    losses.append(eng.train_step().cpu().numpy().copy())
            except RuntimeError as e:
                if mode == "p2p_multicast" and "multicast" in str(e):
                    out["modes"][mode] = {"skipped": str(e)}
                    log(f"{mode}: skipped ({e})")
                    continue
                raise
            torch.cuda.synchronize()
            used_multicast = bool(getattr(eng, "_multicast", False))
            rel = [abs(float(l[2]) - float(r[2])) / abs(float(r[2])) for l, r in zip(losses, ref_losses)]
            parts = [max(abs(float(l[i]) - float(r[i])) / abs(float(r[i])) for l, r in zip(losses[:1], ref_losses[:1]))
                     for i in (0, 1)]
            wdiff = float((model.store.params - ref_params).abs().max().item())
            wmean = float((model.store.params - ref_params).abs().mean().item())
            same = _bits_equal_across_ranks(model.store.params)
            eng.gather_adam_state()
            mdiff = float((model.store.adam_m - ref_m).abs().max().item() / max(ref_m.abs().max().item(), 1e-30))
            vdiff = float((model.store.adam_v - ref_v).abs().max().item() / max(ref_v.abs().max().item(), 1e-30))
            m_same = _bits_equal_across_ranks(model.store.adam_m) and _bits_equal_across_ranks(model.store.adam_v)
            res = {"loss": [float(l[2]) for l in losses], "loss_rel_err": rel, "bce_kl_rel_err_step1": parts,
                   "weights_max_abs_diff": wdiff, "weights_mean_abs_diff": wmean, "replicas_bit_identical": same,
                   "adam_m_max_rel_diff": mdiff, "adam_v_max_rel_diff": vdiff, "adam_state_identical_after_gather": m_same,
                   "multicast": used_multicast, "step_counter": int(model.store.step.item())}
            t = torch.tensor([wdiff, mdiff, vdiff] + rel, dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # worst rank
            res["weights_max_abs_diff"], res["adam_m_max_rel_diff"], res["adam_v_max_rel_diff"] = (float(v) for v in t[:3])
            res["loss_rel_err"] = [float(v) for v in t[3:]]
            out["modes"][mode] = res
            log(f"{mode}: loss_rel_err {res['loss_rel_err']}, weights max diff {wdiff:.3e} (mean {wmean:.3e}), "
                f"replicas identical {same}, multicast {used_multicast}")
            del eng, model
            torch.cuda.empty_cache()
    finally:
        for k, v in saved_env.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    return out


LOSS_TOL_FIRST, LOSS_TOL_LATER, WEIGHT_TOL, ADAM_TOL = 1e-6, 1e-4, 2.5e-3, 1e-3


def verdict(out):
    """List of violated bars (empty = pass)."""
    bad = []
    for mode, r in out["modes"].items():
        if "skipped" in r:
            continue
        if r["loss_rel_err"][0] > LOSS_TOL_FIRST:
            bad.append(f"{mode}: first-step loss differs from the 1-GPU step by {r['loss_rel_err'][0]:.2e} (> {LOSS_TOL_FIRST})")
        if max(r["loss_rel_err"]) > LOSS_TOL_LATER:
            bad.append(f"{mode}: loss differs by {max(r['loss_rel_err']):.2e} (> {LOSS_TOL_LATER})")
        if r["weights_max_abs_diff"] > WEIGHT_TOL:
            bad.append(f"{mode}: weights differ by {r['weights_max_abs_diff']:.2e} (> {WEIGHT_TOL})")
        if not r["replicas_bit_identical"]:
            bad.append(f"{mode}: replicas are not bit-identical across ranks")
        if not r["adam_state_identical_after_gather"] or max(r["adam_m_max_rel_diff"], r["adam_v_max_rel_diff"]) > ADAM_TOL:
            bad.append(f"{mode}: Adam state after gather differs (m {r['adam_m_max_rel_diff']:.2e}, v {r['adam_v_max_rel_diff']:.2e})")
        if r["step_counter"] != out["steps"]:
            bad.append(f"{mode}: step counter {r['step_counter']} != {out['steps']}")
    return bad


def main(argv=None):
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--modes", default="p2p_unicast,p2p_multicast,nccl")
    ap.add_argument("--num-cards", type=int, default=None)
    ap.add_argument("--global-batch", type=int, default=None)
    args = ap.parse_args(argv)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)                                   # NCCL banners etc. go to stderr; stdout carries the JSON line
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank = dist.get_rank()
    log = (lambda m: print(f"[dp_check] {m}", file=sys.stderr, flush=True)) if rank == 0 else (lambda m: None)
    out = run_check(args.precision, args.steps, tuple(args.modes.split(",")), args.num_cards, args.global_batch, log)
    out["violations"] = verdict(out)
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()
    return 1 if out["violations"] else 0


if __name__ == "__main__":
    sys.exit(main())
