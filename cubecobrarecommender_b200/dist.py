"""Host-side data-parallel plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in
the CPU tests).  SURVEY.md §8e: the path shards by cubes with one exchange step per phase --

* graph build: every rank counts its own cube shard, ``all_reduce(SUM)`` of the int32 counts (exact);
* train step : batch and regulariser rows split B/G and R/G per rank, each rank scales its loss
  terms by the GLOBAL B*C and R, the gradients are summed over the ranks (``all_reduce(SUM)`` of the flat
  buffer in the NCCL modes; reduce-scatter + Adam + all-gather in one peer-memory kernel in the default p2p
  mode) and the 3 loss scalars all_reduced; the weights are replicated and stay bit-identical across
  ranks.  Adam's m / v are replicated in the NCCL modes and SLICED in p2p mode (a rank only updates the
  slice it owns, ``owner_slice``): ``DAEEngine.gather_adam_state()`` completes them before a checkpoint;
* inference  : cubes are independent, no collective.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) of ``n`` items for ``rank`` (sizes differ by at most one)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_batch_ids(batch_ids: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The slice of a global batch this rank trains on (global batch must divide evenly so every
    rank launches identical kernel shapes)."""
    n = len(batch_ids)
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by {world} ranks")
    per = n // world
    return batch_ids[rank * per:(rank + 1) * per]


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def all_reduce_sum_(tensor, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def loss_scales(global_batch: int, global_reg_rows: int, num_cards: int, reg: float):
    """(1/(B*C), reg/R) with the GLOBAL sizes: summing the per-rank gradients then reproduces the
    single-process gradient of  mean_{b,c} BCE + reg * mean_r KL  (reference train.py:83-88)."""
    return 1.0 / (float(global_batch) * float(num_cards)), (reg / float(global_reg_rows)) if global_reg_rows else 0.0


def owner_slice(total: int, rank: int, world: int):
    """Slice ``[lo, hi)`` of the flat parameter buffer that ``rank`` owns in the peer-memory data-parallel step
    (cc_adam_step_p2p): the slices tile ``[0, total)`` exactly and start on 4-element (16-byte) boundaries."""
    quarter = total // 4
    lo = (quarter * rank // world) * 4
    hi = (quarter * (rank + 1) // world) * 4 if rank + 1 < world else total
    return lo, hi


def bucket_owner_slice(lo: int, hi: int, rank: int, world: int):
    """``owner_slice`` inside the bucket ``[lo, hi)`` of the flat buffer (bucket bounds are 4-element aligned)."""
    a, b = owner_slice(hi - lo, rank, world)
    return lo + a, lo + b


class GradBuckets:
    """Contiguous slices of the flat gradient buffer in the order backward finishes them:
    "main" decoder, "reg" decoder, then the shared "enc"oder (whose 512 x C first-layer gradient is
    the last thing backward produces).  Each bucket is all_reduced asynchronously as soon as it is
    complete, so the exchange of the two decoders (2/3 of all parameters) overlaps the rest of backward,
    and Adam runs bucket by bucket as the reductions land."""

    ORDER = ("main", "reg", "enc")

    def __init__(self, layout: dict, total: int):
        def span(prefix):
            offs = [(off, off + int(np.prod(shape))) for k, (off, shape) in layout.items() if k.startswith(prefix)]
            return min(o for o, _ in offs), max(e for _, e in offs)
        enc, main, reg = span("encoder_"), span("main_"), span("reg_")
        # the flat buffer is laid out encoder | main | reg with 4-float padding between tensors: extend every
        # bucket to the start of the next one so the three slices tile [0, total) exactly
        assert enc[0] == 0 and enc[1] <= main[0] and main[1] <= reg[0] and reg[1] <= total
        self.ranges = {"enc": (0, main[0]), "main": (main[0], reg[0]), "reg": (reg[0], total)}
        self.works = {}

    def slice(self, flat, name):
        lo, hi = self.ranges[name]
        return flat[lo:hi]

    def launch(self, flat, name, group=None):
        import torch.distributed as dist
        self.works[name] = dist.all_reduce(self.slice(flat, name), op=dist.ReduceOp.SUM, group=group, async_op=True)

    def wait(self, name):
        w = self.works.pop(name, None)
        if w is not None:
            w.wait()
