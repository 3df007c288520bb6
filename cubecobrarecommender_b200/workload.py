"""Synthetic workloads named by BASELINE.json `configs` (shared by bench.py, smoke() and tests)."""
from __future__ import annotations


from .sparse import CubeCSR
from .synth import synth_cubes_csr

# configs[1]: "Regularised DAE train step, ml_files/recommender architecture dims, batch 4096,
#              1xB200 fp32" -- C = 20 884 is the card count of the shipped checkpoints
TRAIN_STEP = dict(workload="dae_train_step", num_cards=20884, batch=4096, reg_rows=4096, num_cubes=8192,
                  reg=0.1, noise=0.2, noise_std=0.1, cfg=2)


def make_cubes(num_cubes: int, num_cards: int, cfg: int) -> CubeCSR:
    ip, ix = synth_cubes_csr(num_cubes, num_cards, cfg=cfg)
    return CubeCSR(ip, ix, num_cards)


def train_step_flops(batch: int, reg_rows: int, num_cards: int) -> float:
    """Algorithmic flops of one step: per tower 3 big passes
    (fwd, dX, dW) of 2*rows*512*C and 6 mid-layer passes x 3, sparse first layer excluded."""
    mid = 512 * 256 + 256 * 128 + 128 * 64 + 64 * 128 + 128 * 256 + 256 * 512
    per_row = 3 * 2 * 512 * num_cards + 3 * 2 * mid
    return float(per_row) * (batch + reg_rows)
