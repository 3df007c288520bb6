#!/usr/bin/env python
"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck) over the kernels with shared-memory
atomics, barriers or peer pointers: noise_kernel, cooc_count_kernel (popcount tiles), topn_rowselect_kernel (both forms),
topn_warpselect_kernel, softmax_kl_regs_kernel / softmax_kl_persistent_kernel, bag_fwd / bag_bwd, adam_p2p_kernel (three
ranks emulated on one GPU).  Results are checked against the oracle as in the tests, so a sanitizer-clean run is also a
correct one.

    compute-sanitizer --tool memcheck  python profiles/sanitizer_driver.py > profiles/r02/sanitizer_memcheck.log 2>&1
    compute-sanitizer --tool racecheck python profiles/sanitizer_driver.py > profiles/r02/sanitizer_racecheck.log 2>&1
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cubecobrarecommender_b200 import _lib, graph as G  # noqa: E402
from cubecobrarecommender_b200._lib import call, ptr, stream_ptr  # noqa: E402
from cubecobrarecommender_b200.ml import engine as E, model as M  # noqa: E402
from cubecobrarecommender_b200.sparse import CubeCSR  # noqa: E402
from cubecobrarecommender_b200.synth import synth_cubes_csr  # noqa: E402
from oracle import graph as og  # noqa: E402

torch.cuda.set_device(0)
dev = "cuda"
c, k, b = 384, 96, 32
ip, ix = synth_cubes_csr(k, c, size_lo=12, size_hi=70, seed=7)
csr = CubeCSR(ip, ix, c)
# --- cooc_count_kernel (popcount path) + normalise
gr = G.build_graph(csr, dev, method="popcount")
assert np.array_equal(gr.counts.cpu().numpy(), og.cooc_counts(ip, ix, c))
print("cooc_count_kernel ok")
# --- noise_kernel + reg rows
prob, alias = E.alias_table(gr.neg_sampler.cpu().numpy(), dev)
model = M.CC_Recommender(c, device=dev, precision="fp32")
eng = E.DAEEngine(model, gr.mhat, batch=b, reg_rows=b, reg=0.1, max_cube_size=80)
indptr, indices = G.upload_csr(csr, dev)
eng.sample_batch(indptr, indices, torch.arange(b, dtype=torch.int32, device=dev), prob, alias, seed=3)
eng.check_overflow()
assert int(eng.x_len.min().item()) > 0
print("noise_kernel ok")
# --- bag_fwd / bag_bwd, losses, Adam (exact-fp32 step: SIMT GEMMs only, no tensor-map kernels under the sanitizer)
loss = eng.train_step().cpu().numpy()
assert np.isfinite(loss).all()
print("fp32 train step ok", loss)
# --- softmax-KL persistent kernels
cc, rows = 340, 40
cpad = 384
g = torch.Generator(device=dev).manual_seed(1)
z = torch.randn(rows, cpad, device=dev, generator=g) * 3
z[:, cc:] = 0
t = torch.rand(16, cc, device=dev, generator=g); t = (t / t.sum(1, keepdim=True)).contiguous()
tr = torch.randint(0, 16, (rows,), dtype=torch.int32, device=dev, generator=g)
table = torch.zeros(16, dtype=torch.float64, device=dev)
call("cc_kl_target_table", ptr(t), cc, 16, cc, ptr(table), stream_ptr())
outs = []
for variant in (0, 1):
    call("cc_softmax_kl_set_variant", variant)
    dz = torch.zeros(rows, cpad, device=dev); rl = torch.zeros(rows, dtype=torch.float64, device=dev); db = torch.zeros(cc, device=dev)
    call("cc_softmax_kl_fwd_bwd_ex", ptr(z), cpad, ptr(t), cc, ptr(tr), rows, cc, cpad, 0.1 / rows, ptr(dz), cpad, ptr(rl), 0,
         ptr(db), None, 0, ptr(table), stream_ptr())
    outs.append((dz, rl))
call("cc_softmax_kl_set_variant", 0)
assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-4, atol=1e-10) and torch.allclose(outs[0][1], outs[1][1], rtol=1e-6)
print("softmax_kl kernels ok")
# --- top-N kernels: streaming select, row select (shared-memory sweeps, 1 and 2 CTAs per SM) and its register form
cn, ld, nb, n = 2001, 2004, 40, 50
rng = np.random.default_rng(0)
vals = rng.standard_normal((nb, cn)).astype(np.float32)
vals[1] = np.round(vals[1] * 2) / 2                       # heavy ties: the survivor buffer overflows
full = torch.full((nb, ld), 1e30, device=dev); full[:, :cn] = torch.from_numpy(vals).to(dev)
lists = [np.sort(rng.choice(cn, size=rng.integers(0, 300), replace=False)) for _ in range(nb)]
mp = np.zeros(nb + 1, np.int64); mp[1:] = np.cumsum([len(x) for x in lists])
mpt = torch.from_numpy(mp).to(dev); mit = torch.from_numpy(np.concatenate(lists).astype(np.int32)).to(dev)
ref = None
for algo in (1, 2, 3, 4):
    call("cc_topn_set_algo", algo)
    ids, v, cnt = G.topn_masked(full[:, :cn], mpt, mit, n, sigmoid=True)
    if ref is None:
        ref = ids.clone()
    assert torch.equal(ids, ref), algo
call("cc_topn_set_algo", 0)
print("top-N kernels ok")
# --- adam_p2p_kernel on three emulated ranks
world, nn = 3, 4 * 1001
p0 = torch.randn(nn, device=dev, generator=g)
params = [p0.clone() for _ in range(world)]; grads = [torch.randn(nn, device=dev, generator=g) * 0.01 for _ in range(world)]
ms = [torch.zeros(nn, device=dev) for _ in range(world)]; vs = [torch.zeros(nn, device=dev) for _ in range(world)]
step = torch.zeros(1, dtype=torch.int64, device=dev)
gp = np.array([x.data_ptr() for x in grads], dtype=np.uint64); pp = np.array([x.data_ptr() for x in params], dtype=np.uint64)
q = nn // 4
bounds = [(q * r // world) * 4 for r in range(world)] + [nn]
for r in range(world):
    call("cc_adam_step_p2p", ptr(gp), ptr(pp), world, r, ptr(ms[r]), ptr(vs[r]), bounds[r], bounds[r + 1], ptr(step), 1e-3, 0.9,
         0.999, 1e-7, None, None, stream_ptr())
torch.cuda.synchronize()
assert torch.equal(params[0], params[1]) and torch.equal(params[0], params[2]) and not torch.equal(params[0], p0)
print("adam_p2p_kernel ok")
print("sanitizer driver finished")
