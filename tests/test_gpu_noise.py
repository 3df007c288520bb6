"""CUDA noise kernel (Philox + alias table) vs the oracle DataGenerator: the invariants of F
and its distribution (MT19937 draws cannot be reproduced, SURVEY.md §5 "Seeding")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from cubecobrarecommender_b200.ml import engine as E
from cubecobrarecommender_b200.sparse import CubeCSR
from cubecobrarecommender_b200.synth import csr_to_dense, synth_cubes_csr
from oracle import graph as og, noise as on


def _run_noise(csr, ns, steps, noise=0.2, std=0.1, seed=123, max_size=128):
    c, b = csr.num_cards, csr.num_cubes
    prob, alias = E.alias_table(ns, "cuda")
    indptr = torch.from_numpy(csr.indptr).cuda(); indices = torch.from_numpy(csr.indices).cuda()
    x_stride = int(max_size * 1.8) + 8
    yw = (c + 127) // 128 * 128 // 32
    x_idx = torch.zeros((b, x_stride), dtype=torch.int32, device="cuda")
    x_len = torch.zeros(b, dtype=torch.int32, device="cuda")
    yb = torch.zeros((b, yw), dtype=torch.int32, device="cuda")
    flips = torch.zeros(b, dtype=torch.int32, device="cuda")
    ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    xs, ys, fs = [], [], []
    for _ in range(steps):
        E.call("cc_noise", E.ptr(indptr), E.ptr(indices), None, b, c, E.ptr(prob), E.ptr(alias), noise, std, seed,
               E.ptr(step), max_size, x_stride, E.ptr(x_idx), E.ptr(x_len), E.ptr(yb), yw, E.ptr(flips), E.ptr(ovf),
               None, 0, E.stream_ptr())
        E.call("cc_step_increment", E.ptr(step), E.stream_ptr())
        xi, xl = x_idx.cpu().numpy(), x_len.cpu().numpy()
        x = np.zeros((b, c), dtype=np.int8)
        for r in range(b):
            row = xi[r, :xl[r]]
            assert len(np.unique(row)) == len(row)              # no duplicates in the x list
            x[r, row] = 1
        bits = yb.cpu().numpy().view(np.uint32)
        y = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(b, -1)[:, :c].astype(np.int8)
        xs.append(x); ys.append(y); fs.append(flips.cpu().numpy().copy())
    assert int(ovf.item()) == 0
    return np.stack(xs), np.stack(ys), np.stack(fs)


def test_noise_invariants_and_distribution():
    c, k = 400, 48
    ip, ix = synth_cubes_csr(k, c, size_lo=20, size_hi=120, seed=21)
    csr = CubeCSR(ip, ix, c)
    dense = csr_to_dense(ip, ix, c)
    mh = og.m_hat(og.create_adjacency_matrix(dense))
    ns = og.neg_sampler(mh)
    steps = 150
    xs, ys, fs = _run_noise(csr, ns, steps)
    size = dense.sum(1)
    for t in range(steps):
        on.check_noise_invariants(dense, xs[t], ys[t])
        assert (fs[t] >= np.floor(0.05 * size - 1e-9)).all() and (fs[t] <= np.floor(0.8 * size)).all()
    # two steps differ, same step reproduces
    assert (xs[0] != xs[1]).any()
    xs2, _, _ = _run_noise(csr, ns, 2)
    assert np.array_equal(xs2, xs[:2])
    # --- flip amount: E[int(s*clip(N(.2,.1),.05,.8))]
    rng = np.random.default_rng(0)
    nz = np.clip(rng.normal(0.2, 0.1, size=200000), 0.05, 0.8)
    expect_flip = np.array([np.floor(s * nz).mean() for s in size])
    got_flip = fs.mean(0)
    assert np.abs(got_flip - expect_flip).max() < 0.08 * expect_flip.max() + 0.5
    # --- distinct removed: P(position removed | f) = 1-(1-1/s)^f ; added: 1-(1-p_c)^f
    removed = ((dense[None] == 1) & (xs == 0)).sum(2)          # (steps, k)
    exp_removed = size[None] * (1 - (1 - 1 / size[None]) ** fs)
    assert abs(removed.sum() - exp_removed.sum()) < 0.02 * exp_removed.sum()
    added = (dense[None] == 0) & (xs == 1)                      # (steps, k, c)
    pex = np.where(dense == 0, ns[None], 0.0); pex /= pex.sum(1, keepdims=True)
    exp_added = 1 - (1 - pex[None]) ** fs[:, :, None]
    # per-card frequency over all steps and cubes follows neg_sampler restricted to the excludes
    got_c, exp_c = added.sum((0, 1)), exp_added.sum((0, 1))
    sd = np.sqrt(exp_c + 1)
    assert (np.abs(got_c - exp_c) < 6 * sd).all()
    assert abs(got_c.sum() - exp_c.sum()) < 0.02 * exp_c.sum()
    # --- y: removed-from-y within removed-from-x, count <= flip//4, mean tracks the oracle generator
    yrem = ((dense[None] == 1) & (ys == 0)).sum(2)
    assert (yrem <= fs // 4).all()
    np.random.seed(5)
    gen = on.DataGenerator(mh, dense, batch_size=k, shuffle=False, noise=0.2)
    o_yrem, o_rem, o_add = [], [], []
    for _ in range(40):
        (xo, _), (yo, _) = gen[0]
        o_yrem.append(((dense == 1) & (yo == 0)).sum()); o_rem.append(((dense == 1) & (xo == 0)).sum())
        o_add.append(((dense == 0) & (xo == 1)).sum())
    assert abs(yrem.sum(1).mean() - np.mean(o_yrem)) < 0.06 * np.mean(o_yrem)
    assert abs(removed.sum(1).mean() - np.mean(o_rem)) < 0.05 * np.mean(o_rem)
    assert abs(added.sum((1, 2)).mean() - np.mean(o_add)) < 0.05 * np.mean(o_add)


def test_reg_rows_follow_neg_sampler():
    c = 300
    rng = np.random.default_rng(2)
    ns = rng.random(c) ** 3; ns[17] = 0; ns /= ns.sum()
    prob, alias = E.alias_table(ns, "cuda")
    n = 400000
    rows = torch.zeros(n, dtype=torch.int32, device="cuda")
    E.call("cc_sample_reg_rows", E.ptr(prob), E.ptr(alias), c, n, 99, None, E.ptr(rows), E.stream_ptr())
    cnt = np.bincount(rows.cpu().numpy(), minlength=c)
    assert cnt[17] == 0 and cnt.sum() == n
    sd = np.sqrt(n * ns * (1 - ns)) + 1
    assert (np.abs(cnt - n * ns) < 6 * sd).all()
