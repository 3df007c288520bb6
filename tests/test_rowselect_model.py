"""The selection algorithm of the CTA-per-cube row select (csrc/topn.cu, topn_rowselect_kernel), restated on the host in
tests/rowselect_model.py, against numpy's stable argsort under the documented tie rule (descending: larger index first,
ascending: smaller index first).  The CUDA kernel itself is compared with the other two select kernels and with numpy in
tests/test_gpu_graph.py; this test pins the ARGUMENT the kernel relies on: the n-th largest thread leader (taken on
32-bit stand-ins, low bits cleared) is a valid lower bound of the answer, and raising it after an overflow terminates."""
import numpy as np
import pytest

from rowselect_model import NT, expect, rowselect


def _row(kind, c, rng):
    if kind == "normal":
        return rng.standard_normal(c).astype(np.float32)
    if kind == "ties":
        return (rng.integers(0, 30, c) / 4 - 3).astype(np.float32)
    if kind == "ascending":
        return np.sort(rng.standard_normal(c).astype(np.float32))
    if kind == "descending":
        return np.sort(rng.standard_normal(c).astype(np.float32))[::-1].copy()
    if kind == "few_strides":                  # the large values sit in 40 of the 512 thread strides
        row = rng.standard_normal(c).astype(np.float32)
        row[((np.arange(c) // 4) % NT) < 40] += 100
        return row
    row = rng.standard_normal(c).astype(np.float32)          # infinities and signed zeros
    row[rng.random(c) < 0.7] = -np.inf if c % 2 else np.inf
    row[rng.random(c) < 0.05] = -0.0
    row[rng.random(c) < 0.05] = 0.0
    return row


@pytest.mark.parametrize("kind", ["normal", "ties", "ascending", "descending", "few_strides", "infinities"])
@pytest.mark.parametrize("c,n", [(97, 7), (700, 50), (5001, 128), (5000, 1)])
def test_rowselect_model_equals_stable_argsort(kind, c, n):
    rng = np.random.default_rng(c * 131 + n)
    row = _row(kind, c, rng)
    listed = [int(i) for i in rng.choice(c, size=int(rng.integers(0, min(c, 700))), replace=False)]
    listed += listed[:3] + [-1, c, c + 5]                     # duplicates and out-of-range entries are ignored
    for only_listed, desc in ((False, True), (True, False), (False, False), (True, True)):
        for cap in (1024, 160):                               # 160: forces the overflow / re-sweep path
            got, cnt, rounds, m = rowselect(row, listed, only_listed, desc, n, cap)
            exp = expect(row, listed, only_listed, desc, n)
            assert got == exp and cnt == len(exp), (kind, c, n, only_listed, desc, cap)


def test_rowselect_model_few_candidates_and_survivor_count():
    rng = np.random.default_rng(5)
    c, n = 20884, 50
    row = (rng.standard_normal(c) * 3 - 4).astype(np.float32)
    listed = [int(i) for i in rng.choice(c, size=540, replace=False)]
    got, cnt, rounds, m = rowselect(row, listed, False, True, n)
    assert got == expect(row, listed, False, True, n) and rounds == 0
    assert n <= m <= 2 * n                    # ~1.3 n elements lie above the leaders' threshold
    got, cnt, _, _ = rowselect(row, list(range(3, c)), False, True, n)          # only three candidates are left
    assert cnt == 3 and got == expect(row, list(range(3, c)), False, True, n)


def _sigmoid_f32(z):
    f = np.float32
    with np.errstate(over="ignore"):
        return (f(1) / (f(1) + np.exp(-z.astype(f)).astype(f))).astype(f)


def _raw_bound_sigmoid(p, descending):
    """rs_raw_bound<SIGMOID=true> of csrc/topn.cu in float32 arithmetic: the logit of the threshold probability,
    moved 2e-6 relative in probability and 1e-5 (1 + |L|) + 2.5e-7 / (1 - p) in the logit to the safe side."""
    f = np.float32
    p = p.astype(f)
    with np.errstate(all="ignore"):
        if descending:
            pm = (p * f(1 - 2e-6)).astype(f) - f(1e-37)
            lg = np.log((pm / (f(1) - pm)).astype(f)).astype(f)
            zb = (lg - (f(1e-5) * (f(1) + np.abs(lg)) + f(2.5e-7) / (f(1) - pm)).astype(f)).astype(f)
            return np.where(pm <= 0, -np.inf, zb).astype(f)
        pp = (p * f(1 + 2e-6)).astype(f) + f(1e-37)
        lg = np.log((pp / (f(1) - pp)).astype(f)).astype(f)
        zb = (lg + (f(1e-5) * (f(1) + np.abs(lg)) + f(2.5e-7) / (f(1) - pp)).astype(f)).astype(f)
        return np.where(pp >= 1, np.inf, zb).astype(f)


def test_fused_sigmoid_logit_bound_never_rejects_a_qualifying_element():
    """The pre-filter of the fused-sigmoid select compares raw logits with a float32 bound derived from the threshold
    probability.  It must be safe: every logit whose float32 sigmoid is at least (descending) / at most (ascending) the
    threshold probability passes the bound -- including the plateaus where the sigmoid saturates to exactly 1.0 / 0.0."""
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.uniform(-110, 20, 1_500_000), rng.uniform(10, 18, 500_000), rng.uniform(-20, -5, 500_000),
                        rng.uniform(-1, 1, 1_000_000), rng.uniform(-0.05, 0.05, 500_000), rng.uniform(3, 12, 500_000),
                        np.array([-np.inf, -104.0, -88.0, 0.0, 16.0, 16.7, 17.0, 30.0, np.inf])]).astype(np.float32)
    order = np.argsort(z, kind="stable")
    zs, ps = z[order], _sigmoid_f32(z[order])
    # descending: the smallest logit whose probability is >= ps[i] must satisfy z >= bound(ps[i])
    first = np.searchsorted(np.maximum.accumulate(ps), ps, side="left")
    assert not (zs[first] < _raw_bound_sigmoid(ps, True)).any()
    # ascending: the largest logit whose probability is <= ps[i] must satisfy z <= bound(ps[i])
    last = np.searchsorted(np.minimum.accumulate(ps[::-1])[::-1], ps, side="right") - 1
    assert not (zs[last] > _raw_bound_sigmoid(ps, False)).any()
    assert (ps == 1.0).any() and (ps == 0.0).any()          # both saturation plateaus were exercised
