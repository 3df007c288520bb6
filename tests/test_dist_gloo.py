"""world_size-2 gloo tests (CPU) of the data-parallel host logic: cube-sharded
count all_reduce, and per-rank loss scaling + gradient all_reduce == single-process gradient."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cubecobrarecommender_b200 import dist as D
from cubecobrarecommender_b200.sparse import CubeCSR
from cubecobrarecommender_b200.synth import csr_to_dense, synth_cubes_csr
from oracle import dae as od, graph as og


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _counts_job(rank, world):
    ip, ix = synth_cubes_csr(90, 150, size_lo=5, size_hi=40, seed=3)
    csr = CubeCSR(ip, ix, 150)
    sh = csr.shard(rank, world)
    cnt = torch.from_numpy(og.cooc_counts(sh.indptr, sh.indices, 150).astype(np.int32))
    D.all_reduce_sum_(cnt)
    return cnt.numpy()


def test_sharded_counts_allreduce_equals_whole():
    out = _spawn(_counts_job)
    ip, ix = synth_cubes_csr(90, 150, size_lo=5, size_hi=40, seed=3)
    whole = og.cooc_counts(ip, ix, 150)
    for r in out:
        assert np.array_equal(r, whole)


def _problem():
    c, k, b = 64, 40, 8
    ip, ix = synth_cubes_csr(k, c, size_lo=4, size_hi=20, seed=9)
    dense = csr_to_dense(ip, ix, c)
    mh = og.m_hat(og.create_adjacency_matrix(dense))
    rng = np.random.default_rng(1)
    x = dense[:b].copy(); y = dense[:b].copy()
    for i in range(b):
        inc = np.where(x[i] == 1)[0]
        x[i, inc[0]] = 0
    rows = rng.integers(0, c, size=b)
    return c, x, y, rows, mh


def _grad_job(rank, world):
    c, x, y, rows, mh = _problem()
    params = od.init_params(c, seed=0)
    b = len(x)
    ids = D.shard_batch_ids(np.arange(b), rank, world)
    inv_bc, kl_scale = D.loss_scales(b, b, c, 0.1)
    m = od.TorchDAE(params, dtype=torch.float64)
    xt, yt = torch.tensor(x[ids]), torch.tensor(y[ids])
    rt, tt = torch.tensor(rows[ids]), torch.tensor(mh[rows[ids]])
    # per-rank loss with GLOBAL normalisation: sum (not mean) of the local terms times the global scales
    z1 = m.tower(xt, "main")
    w1 = m.p["encoder_e1/kernel"]
    eye_rows = torch.zeros(len(ids), c, dtype=torch.float64); eye_rows[torch.arange(len(ids)), rt] = 1
    z2 = m.tower(eye_rows, "reg")
    bce = (torch.clamp(z1, min=0) - z1 * yt + torch.log1p(torch.exp(-z1.abs()))).sum() * inv_bc
    q = torch.softmax(z2, 1)
    tc = torch.clamp(tt, 1e-7, 1.0)
    kl = (tc * torch.log(tc / torch.clamp(q, 1e-7, 1.0))).sum() * kl_scale
    (bce + kl).backward()
    flat = torch.cat([m.p[k].grad.reshape(-1) for k in sorted(m.p)])
    loss = torch.tensor([float(bce + kl)], dtype=torch.float64)
    D.all_reduce_sum_(flat); D.all_reduce_sum_(loss)
    return flat.numpy(), float(loss)


def test_dp_gradients_equal_single_process():
    out = _spawn(_grad_job)
    c, x, y, rows, mh = _problem()
    params = od.init_params(c, seed=0)
    (tot, _, _), grads = od.loss_and_grads_np(params, x, y, rows, mh[rows], 0.1)
    ref = np.concatenate([grads[k].reshape(-1) for k in sorted(grads)])
    for flat, loss in out:
        assert abs(loss - tot) < 1e-12
        assert np.abs(flat - ref).max() < 1e-12
    assert np.array_equal(out[0][0], out[1][0])           # replicas stay bit-identical


def test_shard_helpers():
    assert [D.shard_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    with pytest.raises(ValueError):
        D.shard_batch_ids(np.arange(7), 0, 2)
    csr = CubeCSR.from_lists([[0], [1, 2], [], [3]], 5)
    parts = [csr.shard(r, 2) for r in range(2)]
    assert sum(p.num_cubes for p in parts) == 4 and parts[1].indices.tolist() == [3]


def _bucket_job(rank, world):
    from cubecobrarecommender_b200.ml.model import ParamStore
    store = ParamStore(64, "cpu")
    bk = D.GradBuckets(store.layout, store.total)
    g = torch.Generator().manual_seed(rank)
    store.grads.copy_(torch.randn(store.total, generator=g))
    whole = store.grads.clone()
    for name in bk.ORDER:                       # asynchronous per-bucket reductions, in backward's order
        bk.launch(store.grads, name)
    for name in bk.ORDER:
        bk.wait(name)
    D.all_reduce_sum_(whole)
    return store.grads.numpy(), whole.numpy(), bk.ranges, store.total


def test_gradient_buckets_tile_the_flat_buffer_and_match_one_allreduce():
    out = _spawn(_bucket_job)
    for bucketed, whole, ranges, total in out:
        assert np.array_equal(bucketed, whole)
        spans = sorted(ranges.values())
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # no gap, no overlap
        assert all(lo % 4 == 0 for lo, _ in spans)                          # 16-byte aligned for the Adam kernel


def test_owner_slices_and_full_identity_shards_partition_exactly():
    from cubecobrarecommender_b200.ml.engine import full_identity_shard
    from cubecobrarecommender_b200.ml.model import ParamStore
    total = ParamStore(20884, "cpu").total
    for world in (1, 2, 3, 4, 8, 16):
        spans = [D.owner_slice(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in spans)           # float4 / 16-byte aligned
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 4 + total % 4
        rows = [full_identity_shard(20884, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 20884 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
