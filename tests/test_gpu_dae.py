"""CUDA DAE kernels and train step vs the oracle (fixed weights, fixed noise output), through
the C ABI.  Tolerances: fp32 path loss <= 1e-5 rel (north-star bar 1e-3), grads <= 2e-4 of max."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from cubecobrarecommender_b200.ml import engine as E, model as M
from cubecobrarecommender_b200.sparse import CubeCSR
from cubecobrarecommender_b200.synth import csr_to_dense, synth_cubes_csr
from oracle import dae as od, graph as og


def _bits_from_dense(y):
    b, c = y.shape
    w = (c + 127) // 128 * 128 // 32
    out = np.zeros((b, w), dtype=np.uint32)
    rr, cc = np.nonzero(y == 1)
    np.bitwise_or.at(out, (rr, cc // 32), (np.uint32(1) << (cc % 32).astype(np.uint32)))
    return out.view(np.int32)


def _problem(c=300, k=200, b=64, r=48, seed=11):
    ip, ix = synth_cubes_csr(k, c, size_lo=10, size_hi=60, seed=seed)
    dense = csr_to_dense(ip, ix, c)
    mh = og.m_hat(og.create_adjacency_matrix(dense))
    rng = np.random.default_rng(seed)
    x = dense[:b].copy(); y = dense[:b].copy()
    for i in range(b):
        inc = np.where(x[i] == 1)[0]; exc = np.where(x[i] == 0)[0]
        cut = rng.choice(inc, size=max(1, len(inc) // 5), replace=False)
        x[i, cut] = 0; y[i, cut[:len(cut) // 4]] = 0
        x[i, rng.choice(exc, size=len(cut), replace=False)] = 1
    rows = rng.integers(0, c, size=r)
    return c, x, y, rows, mh


@pytest.mark.parametrize("ta,tb,m,n,k", [(0, 0, 130, 70, 33), (0, 1, 64, 257, 100), (1, 0, 129, 128, 300),
                                         (1, 1, 17, 19, 23), (0, 0, 256, 512, 64)])
def test_gemm_simt_all_layouts(ta, tb, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m * n + k)
    a = torch.randn((k, m) if ta else (m, k), device="cuda", generator=g)
    b = torch.randn((n, k) if tb else (k, n), device="cuda", generator=g)
    bias = torch.randn(n, device="cuda", generator=g)
    mask = torch.randn(m, n, device="cuda", generator=g)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + bias.double()
    ref = torch.relu(ref) * (mask > 0)
    c = torch.full((m, n), 7.0, device="cuda")
    M.gemm(a, b, c, transa=bool(ta), transb=bool(tb), bias=bias, relu=True, mask=mask)
    assert torch.allclose(c.double(), ref, atol=1e-4, rtol=1e-5)
    c2 = torch.ones((m, n), device="cuda")
    M.gemm(a, b, c2, transa=bool(ta), transb=bool(tb), accumulate=True)
    ref2 = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + 1
    assert torch.allclose(c2.double(), ref2, atol=1e-4, rtol=1e-5)


def test_bag_fwd_bwd_vs_dense():
    c, h, b = 700, 512, 37
    g = torch.Generator(device="cuda").manual_seed(3)
    w = torch.randn(c, h, device="cuda", generator=g)
    bias = torch.randn(h, device="cuda", generator=g)
    ip, ix = synth_cubes_csr(b, c, size_lo=0, size_hi=300, seed=4)
    ip[1] = ip[0]                      # noqa: an empty cube row is legal
    csr = CubeCSR(ip, ix, c)
    # rebuild a valid CSR with an empty first cube
    lists = [ix[ip[i]:ip[i + 1]] for i in range(b)]; lists[0] = np.zeros(0, np.int32)
    csr = CubeCSR.from_lists(lists, c)
    sb = M.SparseBatch.from_csr(csr, "cuda")
    out = torch.empty(b, h, device="cuda")
    M.bag_fwd(w, sb.idx, sb.row_start, sb.row_len, bias, out, relu=True)
    x = torch.tensor(csr.to_dense(np.float32)).cuda()
    ref = torch.relu(x.double() @ w.double() + bias.double())
    assert torch.allclose(out.double(), ref, atol=2e-4, rtol=1e-5)
    gout = torch.randn(b, h, device="cuda", generator=g) * (torch.rand(b, h, device="cuda", generator=g) > 0.5)
    dw = torch.zeros(c, h, device="cuda")
    M.bag_bwd(gout, sb.idx, sb.row_start, sb.row_len, dw)
    assert torch.allclose(dw.double(), x.double().t() @ gout.double(), atol=2e-4, rtol=1e-5)


def test_loss_kernels_vs_oracle():
    c, b = 333, 21
    cpad = (c + 127) // 128 * 128
    rng = np.random.default_rng(0)
    z = (rng.standard_normal((b, c)) * 4).astype(np.float32); z[0, :5] = [60, -60, 0, 100, -100]
    y = (rng.random((b, c)) < 0.1).astype(np.float64)
    zt = torch.zeros(b, cpad, device="cuda"); zt[:, :c] = torch.tensor(z).cuda()
    dz = torch.full((b, cpad), 9.0, device="cuda")
    rl = torch.zeros(b, dtype=torch.float64, device="cuda")
    yb = torch.tensor(_bits_from_dense(y)).cuda()
    E.call("cc_bce_logits_fwd_bwd", E.ptr(zt), cpad, E.ptr(yb), yb.shape[1], b, c, cpad, float(b * c), E.ptr(dz), cpad,
           E.ptr(rl), E.stream_ptr())
    z64 = z.astype(np.float64)
    assert abs(rl.sum().item() / (b * c) - od.bce_from_logits_np(z64, y)) < 1e-6
    ref_dz = (1 / (1 + np.exp(-z64)) - y) / (b * c)
    assert np.abs(dz[:, :c].cpu().numpy() - ref_dz).max() < 1e-9
    assert (dz[:, c:] == 0).all()
    # softmax-KL with clipped entries (a huge logit drives most q below 1e-7)
    z2 = (rng.standard_normal((b, c)) * 3).astype(np.float32); z2[3, 7] = 40.0
    t = rng.random((c, c)); t[t < 0.7] = 0; t /= t.sum(1, keepdims=True)
    rows = rng.integers(0, c, size=b).astype(np.int32)
    z2t = torch.zeros(b, cpad, device="cuda"); z2t[:, :c] = torch.tensor(z2).cuda()
    tt = torch.tensor(t.astype(np.float32)).cuda()
    dz2 = torch.full((b, cpad), 9.0, device="cuda")
    E.call("cc_softmax_kl_fwd_bwd", E.ptr(z2t), cpad, E.ptr(tt), c, E.ptr(torch.tensor(rows).cuda()), b, c, cpad,
           0.1 / b, E.ptr(dz2), cpad, E.ptr(rl), 0, None, E.stream_ptr())
    t32 = t.astype(np.float32).astype(np.float64)[rows]
    q = od.softmax_np(z2.astype(np.float64))
    assert abs(rl.sum().item() / b - od.kld_np(t32, q)) < 2e-6 * abs(od.kld_np(t32, q))
    tc = np.clip(t32, 1e-7, 1); un = (q >= 1e-7) & (q <= 1)
    ref = 0.1 / b * (q * (tc * un).sum(1, keepdims=True) - tc * un)
    got = dz2[:, :c].cpu().numpy()
    # entries whose q sits within float rounding of the 1e-7 clip may flip sides
    edge = np.abs(q - 1e-7) < 1e-12
    assert np.abs(got - ref)[~edge].max() < 1e-8
    assert (dz2[:, c:] == 0).all()
    # persistent form (one CTA per SM walks the rows, next row prefetched): same dz and loss, plus the column sums of
    # dz = the layer's bias gradient; exact (expf/logf) and fast (MUFU) variants, more rows than CTAs
    from cubecobrarecommender_b200 import _lib
    if _lib.load().cc_softmax_kl_fuses_dbias(c, cpad, cpad, c, cpad):
        for fast in (0, 1):
            reps = 7                                        # 7 * b rows > 148 CTAs: every CTA loops
            zr = z2t.repeat(reps, 1).contiguous(); rr = torch.tensor(np.tile(rows, reps)).cuda()
            dz_a = torch.full((reps * b, cpad), 9.0, device="cuda"); dz_b = torch.full((reps * b, cpad), 9.0, device="cuda")
            rl_a = torch.zeros(reps * b, dtype=torch.float64, device="cuda"); rl_b = torch.zeros_like(rl_a)
            db = torch.full((c,), 3.0, device="cuda")
            E.call("cc_softmax_kl_fwd_bwd", E.ptr(zr), cpad, E.ptr(tt), c, E.ptr(rr), reps * b, c, cpad, 0.1 / b,
                   E.ptr(dz_a), cpad, E.ptr(rl_a), fast, None, E.stream_ptr())
            E.call("cc_softmax_kl_fwd_bwd", E.ptr(zr), cpad, E.ptr(tt), c, E.ptr(rr), reps * b, c, cpad, 0.1 / b,
                   E.ptr(dz_b), cpad, E.ptr(rl_b), fast, E.ptr(db), E.stream_ptr())
            # same formulas, different reduction trees (1024 vs 512 threads): equal to rounding
            assert torch.allclose(dz_a, dz_b, rtol=2e-5, atol=1e-12) and torch.allclose(rl_a, rl_b, rtol=1e-6)
            assert (dz_b[:, c:] == 0).all()
            ref_db = dz_b[:, :c].double().sum(0)
            assert (db.double() - ref_db).abs().max().item() < 1e-6 * ref_db.abs().max().item() + 1e-12
            # with the per-target-row table sum t' log t' (cc_kl_target_table) the kernel evaluates no logarithm per
            # element: same dlogits bit for bit, row losses equal to rounding (and closer to the float64 oracle)
            table = torch.zeros(c, dtype=torch.float64, device="cuda")
            E.call("cc_kl_target_table", E.ptr(tt), c, c, c, E.ptr(table), E.stream_ptr())
            tcl = np.clip(t.astype(np.float32).astype(np.float64), 1e-7, 1.0)
            assert np.abs(table.cpu().numpy() - (tcl * np.log(tcl)).sum(1)).max() < 1e-12
            dz_c = torch.full((reps * b, cpad), 9.0, device="cuda"); rl_c = torch.zeros_like(rl_a)
            db_c = torch.zeros(c, device="cuda")
            E.call("cc_softmax_kl_fwd_bwd_ex", E.ptr(zr), cpad, E.ptr(tt), c, E.ptr(rr), reps * b, c, cpad, 0.1 / b,
                   E.ptr(dz_c), cpad, E.ptr(rl_c), fast, E.ptr(db_c), None, 0, E.ptr(table), E.stream_ptr())
            assert torch.equal(dz_c, dz_b)
            assert torch.allclose(rl_c, rl_b, rtol=2e-6)
            kl_ref = od.kld_np(t32, q)
            assert abs(rl_c[:b].sum().item() / b - kl_ref) <= abs(rl_b[:b].sum().item() / b - kl_ref) + 1e-7 * abs(kl_ref)


def test_adam_matches_tf_style_oracle():
    n = 1003
    rng = np.random.default_rng(1)
    p = rng.standard_normal(n).astype(np.float32); g = rng.standard_normal(n).astype(np.float32)
    pt, gt = torch.tensor(p).cuda(), torch.tensor(g).cuda()
    buf = torch.zeros(1008, device="cuda"); buf[:n] = pt
    gb = torch.zeros(1008, device="cuda"); gb[:n] = gt
    m = torch.zeros(1008, device="cuda"); v = torch.zeros(1008, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    P = {"w": p.astype(np.float64)}; Mo = {"w": np.zeros(n)}; Vo = {"w": np.zeros(n)}
    for t in range(1, 4):
        E.call("cc_adam_step", E.ptr(buf), E.ptr(gb), E.ptr(m), E.ptr(v), n, E.ptr(step), 1e-3, 0.9, 0.999, 1e-7,
               None, E.stream_ptr())
        E.call("cc_step_increment", E.ptr(step), E.stream_ptr())
        od.adam_step_np(P, {"w": g.astype(np.float64)}, Mo, Vo, t)
        assert np.abs(buf[:n].cpu().numpy() - P["w"]).max() < 5e-7
    assert int(step.item()) == 3 and (buf[n:] == 0).all()


def _rn_tf32(x):
    xi = x.contiguous().view(torch.int32)
    return ((xi + 0x0FFF + ((xi >> 13) & 1)) & ~0x1FFF).view(torch.float32)


def test_adam_p2p_emulated_ranks_equal_plain_adam():
    """cc_adam_step_p2p (reduce-scatter + Adam + all-gather over peer pointers) with three ranks emulated on ONE
    GPU -- three gradient and three parameter buffers on the same device, one launch per rank slice: every
    rank's parameters must equal plain Adam on the summed gradient, bit for bit, over three steps."""
    world, n = 3, 4 * 3001
    g = torch.Generator(device="cuda").manual_seed(2)
    p0 = torch.randn(n, device="cuda", generator=g)
    params = [p0.clone() for _ in range(world)]
    grads = [torch.empty(n, device="cuda") for _ in range(world)]
    ms = [torch.zeros(n, device="cuda") for _ in range(world)]
    vs = [torch.zeros(n, device="cuda") for _ in range(world)]
    ref_p, ref_m, ref_v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    gp = np.array([t.data_ptr() for t in grads], dtype=np.uint64)
    pp = np.array([t.data_ptr() for t in params], dtype=np.uint64)
    quarter = n // 4
    bounds = [(quarter * r // world) * 4 for r in range(world)] + [n]
    hp = (1e-3, 0.9, 0.999, 1e-7)
    for it in range(3):
        for t in grads:
            t.copy_(torch.randn(n, device="cuda", generator=g) * 0.01)
        total = grads[0].clone()
        for t in grads[1:]:
            total += t                                         # rank order, like the kernel
        for r in range(world):
            E.call("cc_adam_step_p2p", E.ptr(gp), E.ptr(pp), world, r, E.ptr(ms[r]), E.ptr(vs[r]), bounds[r], bounds[r + 1],
                   E.ptr(step), *hp, None, None, E.stream_ptr())
        E.call("cc_adam_step", E.ptr(ref_p), E.ptr(total), E.ptr(ref_m), E.ptr(ref_v), n, E.ptr(step), *hp, None,
               E.stream_ptr())
        E.call("cc_step_increment", E.ptr(step), E.stream_ptr())
        for r in range(world):
            assert torch.equal(params[r], ref_p)               # every rank holds the same, exact update
            lo, hi = bounds[r], bounds[r + 1]
            assert torch.equal(ms[r][lo:hi], ref_m[lo:hi]) and torch.equal(vs[r][lo:hi], ref_v[lo:hi])
            assert ms[r][:lo].abs().sum() == 0 and ms[r][hi:].abs().sum() == 0      # only the own slice is touched


@pytest.fixture
def pair_mode():
    """Sets the CTA-pair (cta_group::2) tiling mode of the tcgen05 GEMMs for one test, then restores 'auto'."""
    from cubecobrarecommender_b200 import _lib

    def set_mode(mode):
        _lib.call("cc_gemm_tc_set_pair_mode", int(mode))
    yield set_mode
    set_mode(-1)


@pytest.mark.parametrize("tile_n", [128, 256, "pair"])
@pytest.mark.parametrize("ta,tb,m,n,k,split", [(0, 0, 130, 72, 132, 1), (0, 1, 64, 260, 100, 1), (1, 0, 132, 128, 300, 1),
                                               (1, 1, 20, 24, 36, 1), (0, 0, 4096, 512, 256, 1), (1, 0, 512, 64, 8192, 0),
                                               (1, 0, 128, 260, 4100, 7), (0, 1, 300, 260, 20884, 1),
                                               (0, 1, 300, 520, 20884, 0)])
def test_gemm_tcgen05_tf32_all_layouts(ta, tb, m, n, k, split, tile_n, pair_mode):
    """tcgen05 kind::tf32 GEMM vs float64 on operands already rounded to tf32 (so the products are exact
    and only fp32 accumulation differs): K-major and MN-major operands, ragged tiles, split-K; single-CTA
    128x128 / 128x256 tiles and 256x256 CTA-pair tiles (cta_group::2)."""
    from cubecobrarecommender_b200.ml import tensorcore as TC
    pair_mode(1 if tile_n == "pair" else 0)
    tile_n = 256 if tile_n == "pair" else tile_n
    g = torch.Generator(device="cuda").manual_seed(m * n + k)
    a = _rn_tf32(torch.randn((k, m) if ta else (m, k), device="cuda", generator=g))
    b = _rn_tf32(torch.randn((n, k) if tb else (k, n), device="cuda", generator=g))
    opa, opb = (a.t() if ta else a).double(), (b.t() if tb else b).double()
    ref = opa @ opb
    scale = ref.abs().max().item()
    c = torch.full((m, n), 7.0, device="cuda")
    TC.gemm(a, b, c, transa=bool(ta), transb=bool(tb), precision="tf32", split_k=split, tile_n=tile_n if split else 0)
    tol = 2e-5 + 4e-9 * k                 # fp32 accumulation in TMEM over k terms
    assert (c.double() - ref).abs().max().item() / scale < tol
    if split == 1:
        bias = torch.randn(n, device="cuda", generator=g)
        mask = torch.randn(m, n, device="cuda", generator=g)
        c2 = torch.full((m, n), 7.0, device="cuda")
        TC.gemm(a, b, c2, transa=bool(ta), transb=bool(tb), bias=bias, relu=True, mask=mask, precision="tf32",
                round_out=True, split_k=1, tile_n=tile_n)
        ref2 = torch.relu(ref + bias.double()) * (mask > 0)
        assert (c2.double() - ref2).abs().max().item() / scale < 6e-4       # output rounded to tf32 (2^-11)
        assert torch.equal(c2, _rn_tf32(c2))
        c3 = torch.ones((m, n), device="cuda")
        TC.gemm(a, b, c3, transa=bool(ta), transb=bool(tb), accumulate=True, precision="tf32", split_k=1,
                tile_n=tile_n)
        assert (c3.double() - ref - 1).abs().max().item() / scale < tol
    if split != 1:      # split-K with a non-linear epilogue = reduce-add pass + elementwise pass
        bias = torch.randn(n, device="cuda", generator=g)
        c4 = torch.full((m, n), 7.0, device="cuda")
        TC.gemm(a, b, c4, transa=bool(ta), transb=bool(tb), bias=bias, relu=True, precision="tf32", split_k=split,
                tile_n=tile_n if split else 0)
        assert (c4.double() - torch.relu(ref + bias.double())).abs().max().item() / scale < tol


@pytest.mark.parametrize("m", [8192, 300])
@pytest.mark.parametrize("widths", [(512, 256, 128, 64), (64, 128, 256, 512), (512, 256), (256, 512), (128, 64, 128)])
@pytest.mark.parametrize("backward", [False, True])
def test_chain_of_small_layers_equals_separate_gemms(widths, m, backward):
    """cc_chain_tc: up to three consecutive Dense layers in one launch with the intermediate activations kept in tensor
    memory (tcgen05.mma with A in TMEM) against the same layers as separate cc_gemm_tc calls -- same k order, same fp32
    accumulation, so every layer's output must be bit-identical.  Forward form (Keras kernels [K][N], bias + ReLU +
    tf32 rounding) and backward form (kernels used transposed, ReLU mask); full row blocks and a ragged last one."""
    from cubecobrarecommender_b200.ml import tensorcore as TC
    g = torch.Generator(device="cuda").manual_seed(sum(widths) + m)
    a = _rn_tf32(torch.randn((m, widths[0] + 4), device="cuda", generator=g))[:, :widths[0]]      # padded row stride
    layers_chain, layers_ref = [], []
    for k, n in zip(widths[:-1], widths[1:]):
        w = _rn_tf32(torch.randn((n, k) if backward else (k, n), device="cuda", generator=g) / k ** 0.5)
        bias = None if backward else torch.randn(n, device="cuda", generator=g) * 0.1
        mask = torch.randn((m, n + 4), device="cuda", generator=g)[:, :n] if backward else None
        out_c = torch.full((m, n + 4), 7.0, device="cuda")[:, :n]
        out_r = torch.full((m, n + 4), 7.0, device="cuda")[:, :n]
        layers_chain.append((w, not backward, bias, mask, out_c))
        layers_ref.append((w, bias, mask, out_r))
    h = a
    for w, bias, mask, out_r in layers_ref:
        TC.gemm(h, w, out_r, transb=backward, bias=bias, relu=not backward, mask=mask, precision="tf32", round_out=True)
        h = out_r
    TC.chain(a, layers_chain, relu=True, round_out=True)
    torch.cuda.synchronize()
    h = a
    for l, ((w, _, bias, mask, out_c), (_, _, _, out_r)) in enumerate(zip(layers_chain, layers_ref)):
        if m == 8192:       # (at 300 rows the planner splits K for the separate GEMMs: another summation order)
            assert torch.equal(out_c, out_r), (l, (out_c - out_r).abs().max().item())
        # layer by layer against float64 on the chain's own input of that layer
        ref = h.double() @ (w.t() if backward else w).double()
        ref = ref * (mask > 0) if backward else torch.relu(ref + bias.double())
        scale = ref.abs().max().item()
        assert scale > 0 and (out_c.double() - ref).abs().max().item() / scale < 6e-4, l      # output rounded to tf32
        assert torch.equal(out_c, _rn_tf32(out_c))
        assert (out_c._base[:, out_c.shape[1]:] == 7.0).all()          # nothing written beyond the layer's width
        h = out_c
    # the ReLU test as one bit per element: written by the forward form, consumed by the backward form
    def pack(t):            # (m, n) bool -> (m, n / 32) int32, bit j of word w = column 32 w + j
        b = t.reshape(t.shape[0], -1, 32).to(torch.int64)
        v = (b << torch.arange(32, device="cuda")).sum(-1)
        return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()
    if backward:
        again = [(w, False, None, pack(mask > 0), torch.full_like(out_c._base, 7.0)[:, :out_c.shape[1]])
                 for (w, _, _, mask, out_c) in layers_chain]
        TC.chain(a, again, round_out=True)
        for l, (one, two) in enumerate(zip(layers_chain, again)):
            assert torch.equal(one[4], two[4]), l
    else:
        bits = [torch.full((m, lay[4].shape[1] // 32), -1, dtype=torch.int32, device="cuda") for lay in layers_chain]
        again = [(w, True, bias, None, torch.full_like(out_c._base, 7.0)[:, :out_c.shape[1]], bt)
                 for (w, _, bias, _, out_c), bt in zip(layers_chain, bits)]
        TC.chain(a, again, relu=True, round_out=True)
        for l, (one, two, bt) in enumerate(zip(layers_chain, again, bits)):
            assert torch.equal(one[4], two[4]) and torch.equal(bt, pack(one[4] > 0)), l


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("ta,tb,m,n,k", [(1, 0, 512, 5000, 4096),      # dW-like: fewer tiles than units, all stream-K
                                         (0, 1, 2500, 512, 20992),     # dX-like: long K, non-linear epilogue (two passes)
                                         (1, 0, 768, 20000, 4096),     # data-parallel waves + a stream-K tail that starts
                                         (1, 1, 1032, 3000, 4104),     # mid-column; ragged M/N/K
                                         (1, 0, 21000, 512, 4096),     # dW1-like: 83 x 2 tiles in column-fastest order
                                         (1, 0, 15296, 600, 4096)])    # column-fastest, stream-K tail starting mid-row
def test_gemm_tcgen05_hybrid_stream_k(ta, tb, m, n, k, pair, precision, pair_mode):
    """Hybrid stream-K schedule (forced on): whole tile waves data-parallel, the rest cut along K into one span per CTA
    (pair) and TMA-reduce-added into pre-zeroed tiles.  Against float64 on exactly representable operands; plain
    store, accumulate, and the non-linear epilogue (bias + ReLU + mask + tf32 rounding as a second pass)."""
    from cubecobrarecommender_b200 import _lib
    from cubecobrarecommender_b200.ml import tensorcore as TC
    pair_mode(pair)
    _lib.call("cc_gemm_tc_set_stream_k", 1)
    try:
        plan = np.zeros(4, dtype=np.int32)
        _lib.call("cc_gemm_tc_plan_ex", 1 if precision == "tf32" else 2, m, n, k, 0, 0, _lib.ptr(plan))
        assert plan[3] == 1 and plan[1] == 1 and plan[2] == 1 + pair
        g = torch.Generator(device="cuda").manual_seed(m + n + k)
        cast = (lambda t: _rn_tf32(t)) if precision == "tf32" else (lambda t: t.to(torch.bfloat16))
        a = cast(torch.randn((k, m) if ta else (m, k), device="cuda", generator=g))
        b = cast(torch.randn((n, k) if tb else (k, n), device="cuda", generator=g))
        ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
        scale = ref.abs().max().item()
        tol = 2e-5 + 4e-9 * k
        c = torch.full((m, n), 7.0, device="cuda")            # garbage: the stream-K tiles must be zeroed by the call
        TC.gemm(a, b, c, transa=bool(ta), transb=bool(tb), precision=precision)
        assert (c.double() - ref).abs().max().item() / scale < tol
        c2 = torch.ones((m, n), device="cuda")
        TC.gemm(a, b, c2, transa=bool(ta), transb=bool(tb), accumulate=True, precision=precision)
        assert (c2.double() - ref - 1).abs().max().item() / scale < tol
        bias = torch.randn(n, device="cuda", generator=g)
        mask = torch.randn(m, n, device="cuda", generator=g)
        c3 = torch.full((m, n), 7.0, device="cuda")
        TC.gemm(a, b, c3, transa=bool(ta), transb=bool(tb), bias=bias, relu=True, mask=mask, precision=precision, round_out=True)
        ref3 = torch.relu(ref + bias.double()) * (mask > 0)
        assert (c3.double() - ref3).abs().max().item() / scale < 6e-4 and torch.equal(c3, _rn_tf32(c3))
        # a padded output (leading dimension > n): the zero fill and the reduce-adds respect the row stride
        buf = torch.full((m, n + 12), 5.0, device="cuda")
        TC.gemm(a, b, buf[:, :n], transa=bool(ta), transb=bool(tb), precision=precision)
        assert (buf[:, :n].double() - ref).abs().max().item() / scale < tol and (buf[:, n:] == 5.0).all()
    finally:
        _lib.call("cc_gemm_tc_set_stream_k", -1)


@pytest.mark.parametrize("pair", [0, 1])
def test_gemm_tcgen05_dynamic_tile_scheduler_many_tiles(pair, pair_mode):
    """Far more tiles than persistent CTAs (or CTA pairs): after its first tile every unit draws the rest from the
    global atomic counter, which must re-arm itself for the next launch (three launches, identical results),
    including while a second stream keeps some SMs busy."""
    from cubecobrarecommender_b200 import _lib
    from cubecobrarecommender_b200.ml import tensorcore as TC
    pair_mode(pair)
    _lib.call("cc_gemm_tc_set_dynamic_tiles", 1)
    try:
        _dynamic_tiles_body(TC)
    finally:
        _lib.call("cc_gemm_tc_set_dynamic_tiles", 0)


def _dynamic_tiles_body(TC):
    m, n, k = 4096 + 64, 6000, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    a = _rn_tf32(torch.randn(m, k, device="cuda", generator=g))
    b = _rn_tf32(torch.randn(k, n, device="cuda", generator=g))
    ref = a.double() @ b.double()
    scale = ref.abs().max().item()
    outs = []
    side = torch.cuda.Stream()
    junk = torch.randn(4096, 4096, device="cuda")
    for it in range(3):
        c = torch.zeros(m, n, device="cuda")
        if it == 2:                      # a competing kernel stream: late-starting units must just take fewer tiles
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(4):
                    junk = torch.tanh(junk @ junk * 1e-3)
        TC.gemm(a, b, c, precision="tf32", split_k=1, tile_n=256)
        outs.append(c)
    torch.cuda.synchronize()
    for c in outs:
        assert (c.double() - ref).abs().max().item() / scale < 3e-5
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("m,c", [(200, 1000), (300, 1100)])
def test_gemm_bce_fused_epilogue_vs_oracle(m, c, pair, pair_mode):
    from cubecobrarecommender_b200.ml import tensorcore as TC
    pair_mode(pair)
    k = 512
    cpad = (c + 127) // 128 * 128
    g = torch.Generator(device="cuda").manual_seed(1)
    a = _rn_tf32(torch.randn(m, k, device="cuda", generator=g))
    w = _rn_tf32(torch.randn(k, c, device="cuda", generator=g) * 0.2)
    bias = torch.randn(c, device="cuda", generator=g)
    y = (torch.rand(m, c, device="cuda", generator=g) < 0.05).double().cpu().numpy()
    yb = torch.tensor(_bits_from_dense(y)).cuda()
    dz = torch.full((m, cpad), 9.0, device="cuda")
    part = torch.zeros(TC.bce_partial_count(m, cpad), dtype=torch.float64, device="cuda")
    dbias = torch.full((c,), 5.0, device="cuda")
    TC.gemm_bce(a, w, bias, yb, float(m * c), dz, part, precision="tf32", round_out=False, dbias=dbias)
    z = (a.double() @ w.double() + bias.double()).cpu().numpy()
    ref_loss = od.bce_from_logits_np(z, y)
    assert abs(part.sum().item() / (m * c) - ref_loss) / ref_loss < 1e-5
    ref_dz = (1 / (1 + np.exp(-z)) - y) / (m * c)
    assert np.abs(dz[:, :c].cpu().numpy() - ref_dz).max() < 2e-6 / (m * c) * 1e3 + 1e-9
    assert (dz[:, c:] == 0).all()
    # bias gradient = column sums of dlogits, reduced by the epilogue (float atomics over 32-row blocks)
    assert np.abs(dbias.cpu().numpy() - ref_dz.sum(0)).max() < 2e-6 * np.abs(ref_dz.sum(0)).max() + 1e-12


# tf32: the north-star bar is the per-step loss (1e-3 relative).  Gradients are held to 1e-2 of their max on
# the first step (identical weights); afterwards Adam's m/sqrt(v) (= sign(g) on step 1) amplifies rounding
# of near-zero gradients into +-lr weight differences, so later steps compare two slightly different nets.
# bf16 (reported separately from the fp32/tf32 headline, BASELINE north star): bf16 operands (8-bit mantissa) in the seven
# 512 <-> C passes, fp32 accumulation and master weights; stated tolerance: loss 2e-3 relative, gradients 5e-2 of their max.
TOL = {"fp32": dict(loss=1e-5, grad=2e-4, grad_later=2e-4, weight=3e-4),
       "tf32": dict(loss=1e-3, grad=1e-2, grad_later=6e-2, weight=2.5e-3),
       "bf16": dict(loss=2e-3, grad=5e-2, grad_later=0.2, weight=4e-3)}


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("r", [48, 0])
def test_train_steps_match_oracle(r, precision, monkeypatch):
    """First layer as the embedding-bag gather over the fp32 master W1 (exact sums): the tolerances of round 1."""
    monkeypatch.setenv("CC_FIRST_LAYER", "gather")
    eng = _train_steps_vs_oracle(r, precision)
    assert not eng.first_layer_tc


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_train_steps_match_oracle_first_layer_on_tensor_cores(precision):
    """The default in the tensor-core modes: the main rows' first layer as a dense x W1 GEMM (W1 enters rounded to tf32 /
    bf16).  Same loss bar; on THIS small problem (cubes of 10-60 cards: no averaging over hundreds of products) the
    first-step gradients are held to 4e-2 (tf32) / 1e-1 (bf16: W1 itself enters with 2^-9 relative error; measured 8e-2
    on encoder_e2/kernel) of their max-norm -- at the BASELINE shape the bars stay 1e-2 / 5e-2
    (tests/test_gpu_baseline_shapes.py)."""
    eng = _train_steps_vs_oracle(48, precision, grad_tol={"tf32": 4e-2, "bf16": 1e-1}[precision])
    assert eng.first_layer_tc


def _train_steps_vs_oracle(r, precision, grad_tol=None):
    c, x, y, rows, mh = _problem(r=max(r, 1))
    rows = rows[:r]
    params = od.init_params(c, seed=0)
    rng = np.random.default_rng(5)
    for kname in params:
        if kname.endswith("bias"):
            params[kname] = (rng.standard_normal(params[kname].shape) * 0.05).astype(np.float32)
    tol = TOL[precision]
    model = M.CC_Recommender(c, device="cuda", precision=precision)
    model.set_weights_dict(params)
    mhat = torch.tensor(mh.astype(np.float32)).cuda()
    eng = E.DAEEngine(model, mhat, batch=x.shape[0], reg_rows=r, reg=0.1, max_cube_size=80)
    sb = M.SparseBatch.from_csr(CubeCSR.from_dense(x), "cuda")
    eng.set_batch(sb, torch.tensor(_bits_from_dense(y)).cuda(), torch.tensor(rows).cuda())
    # oracle in float64 on the float32 targets Keras would see
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    m64 = {k: np.zeros_like(v) for k, v in p64.items()}; v64 = {k: np.zeros_like(v) for k, v in p64.items()}
    t32 = mh.astype(np.float32).astype(np.float64)[rows] if r else np.zeros((0, c))
    for step in range(1, 4):
        if r:
            (tot, bce, kl), grads = od.loss_and_grads_np(p64, x, y, rows, t32, 0.1)
        else:
            (tot, bce, kl), grads = od.loss_and_grads_np(p64, x, y, np.zeros(0, np.int64), np.zeros((0, c)), 0.0)
            tot = bce
        eng.forward_backward()
        got = eng.loss3.cpu().numpy()
        assert abs(got[0] - bce) / bce < tol["loss"]
        if r:
            assert abs(got[1] - kl) / kl < tol["loss"]
            assert abs(got[2] - tot) / tot < tol["loss"]
        gd = model.store.to_dict(model.store.grads)
        for kname, gref in grads.items():
            if not r and kname.startswith("reg_"):
                continue
            scale = np.abs(gref).max() + 1e-30
            bar = tol["grad" if step == 1 else "grad_later"]
            assert np.abs(gd[kname] - gref).max() / scale < max(bar, grad_tol or 0.0), (step, kname)
        eng.apply_adam()
        od.adam_step_np(p64, grads, m64, v64, step)
    pd = model.get_weights_dict()
    for kname, pref in p64.items():
        if not r and kname.startswith("reg_"):
            continue
        assert np.abs(pd[kname] - pref).max() < tol["weight"], kname   # 3 Adam steps move weights by <= 3e-3
    return eng


def test_model_call_api_parity():
    c, x, y, rows, mh = _problem(c=200, b=8, r=5)
    params = od.init_params(c, seed=1)
    model = M.CC_Recommender(c, device="cuda")
    model.set_weights_dict(params)
    eye_rows = np.zeros((5, c)); eye_rows[np.arange(5), rows] = 1
    rec, reg = model((x, eye_rows))
    z1, z2, _ = od.forward_np(params, x, rows)
    assert np.abs(rec.cpu().numpy() - 1 / (1 + np.exp(-z1))).max() < 1e-5
    assert np.abs(reg.cpu().numpy() - od.softmax_np(z2)).max() < 1e-6
    lat = model.encoder(x)
    assert lat.shape == (8, 64)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_host_batch_stream_equals_device_resident_steps(precision):
    """ml.engine.HostBatchStream (pinned host CSR -> double-buffered H2D on a copy stream, loss read back one
    step late) must train exactly like the device-resident path: same seeds, same cubes => identical losses."""
    from cubecobrarecommender_b200 import graph as G
    c, k, b = 384, 96, 32
    ip, ix = synth_cubes_csr(k, c, size_lo=12, size_hi=70, seed=7)
    csr = CubeCSR(ip, ix, c)
    gr = G.build_graph(csr, "cuda", want_m64=False)
    prob, alias = E.alias_table(gr.neg_sampler.cpu().numpy(), "cuda")
    params = od.init_params(c, seed=0)
    losses = []
    for mode in ("device", "host"):
        model = M.CC_Recommender(c, device="cuda", precision=precision)
        model.set_weights_dict(params)
        eng = E.DAEEngine(model, gr.mhat, batch=b, reg_rows=b, reg=0.1, max_cube_size=80)
        out = []
        if mode == "device":      # the same per-step batches, uploaded synchronously (the noise draws are keyed by the
            for i in range(5):    # cube's position in the CSR it is read from, so both paths must see the same CSRs)
                indptr, indices = G.upload_csr(csr.rows(np.arange((i % 3) * b, (i % 3 + 1) * b)), "cuda")
                eng.sample_batch(indptr, indices, None, prob, alias, 0.2, 0.1, seed=11)
                out.append(eng.train_step().cpu().numpy().copy())
        else:
            feed = E.HostBatchStream(eng, [csr.rows(np.arange(j * b, (j + 1) * b)) for j in range(3)])
            for i in range(5):
                l = feed.step(i, prob, alias, 0.2, 0.1, seed=11)
                assert (l is None) == (i == 0)
                if l is not None:
                    out.append(l)
            out.append(feed.drain())
            assert feed.h2d_bytes > 0
        eng.check_overflow()
        losses.append(np.array(out))
    assert losses[0].shape == (5, 3)
    # float atomics (first-layer scatter-add in fp32 mode, bias-gradient column sums in the fused BCE epilogue) make
    # the last bits of a step vary from run to run, so the two runs agree to rounding, not bit for bit
    assert np.array_equal(losses[0][0], losses[1][0])          # first step: identical weights, identical batch
    assert np.allclose(losses[0], losses[1], rtol=1e-5, atol=0)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_full_identity_regulariser_and_its_row_shards(precision, monkeypatch):
    """README form of the regulariser, KL(M-hat, D2(E(I))) over ALL rows of I (BASELINE configs[2]): (a) one engine
    with reg_rows = arange(C) matches the oracle; (b) two engines taking the two row shards, each scaling by the
    GLOBAL row count, produce losses and gradients that sum to (a) -- the data-parallel partition of the KL term.
    The 16 main rows take the gather first layer (exact sums over the fp32 W1): on a 200-card problem a tensor-core
    first layer (W1 rounded to tf32) flips enough ReLUs to move single gradient entries by 4-8% of the max-norm, which
    says nothing about the row partition this test is for; that layer has its own tests and the BASELINE-shape bar."""
    monkeypatch.setenv("CC_FIRST_LAYER", "gather")
    c, x, y, _, mh = _problem(c=200, b=16, r=1)
    params = od.init_params(c, seed=2)
    tol = TOL[precision]
    mhat = torch.tensor(mh.astype(np.float32)).cuda()
    sb = M.SparseBatch.from_csr(CubeCSR.from_dense(x), "cuda")
    yb = torch.tensor(_bits_from_dense(y)).cuda()
    rows_all = np.arange(c)
    t32 = mh.astype(np.float32).astype(np.float64)
    (tot, bce, kl), grads = od.loss_and_grads_np({k: v.astype(np.float64) for k, v in params.items()}, x, y, rows_all, t32, 0.1)

    def run(lo, hi, main_rows):
        model = M.CC_Recommender(c, device="cuda", precision=precision)
        model.set_weights_dict(params)
        eng = E.DAEEngine(model, mhat, batch=x.shape[0], reg_rows=hi - lo, reg=0.1, max_cube_size=80,
                          global_batch=x.shape[0], global_reg_rows=c)
        eng.set_batch(sb, yb, torch.arange(lo, hi, dtype=torch.int32, device="cuda"))
        eng.forward_backward()
        return eng.loss3.cpu().numpy().copy(), model.store.to_dict(model.store.grads)

    l_full, g_full = run(0, c, True)
    assert abs(l_full[1] - kl) / kl < tol["loss"] and abs(l_full[2] - tot) / tot < tol["loss"]
    for kname, gref in grads.items():       # (C reg rows on top of the batch: twice the single-step tf32 allowance)
        assert np.abs(g_full[kname] - gref).max() / (np.abs(gref).max() + 1e-30) < 2 * tol["grad"], kname
    # two row shards (as two ranks would hold them): KL parts add up; BCE is computed by both here, so compare KL only
    lo0, hi0 = E.full_identity_shard(c, 0, 2); lo1, hi1 = E.full_identity_shard(c, 1, 2)
    assert (lo0, hi1) == (0, c) and hi0 == lo1
    l0, g0 = run(lo0, hi0, True); l1, g1 = run(lo1, hi1, True)
    assert abs((l0[1] + l1[1]) - l_full[1]) / l_full[1] < 1e-5
    for kname in ("reg_reconstruction/kernel", "reg_d1/kernel", "reg_reconstruction/bias"):
        ssum = g0[kname] + g1[kname]
        assert np.abs(ssum - g_full[kname]).max() / (np.abs(g_full[kname]).max() + 1e-30) < (1e-5 if precision == "fp32" else 2e-3), kname


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("ta,tb,m,n,k", [(0, 0, 130, 72, 136), (0, 1, 64, 264, 104), (1, 0, 136, 128, 304), (1, 1, 24, 24, 40),
                                         (0, 0, 4096, 512, 256), (1, 0, 512, 264, 4096), (0, 1, 300, 520, 20992)])
def test_gemm_tcgen05_bf16_all_layouts(ta, tb, m, n, k, pair, pair_mode):
    """tcgen05 kind::f16 GEMM on bf16 operands (fp32 accumulation, fp32 output): products of bf16 values are exact in
    fp32, so only the accumulation order differs from the float64 reference.  K-major and MN-major operands, ragged
    tiles, CTA pairs.  (Leading dimensions are multiples of 8 elements: TMA rows are multiples of 16 bytes.)"""
    from cubecobrarecommender_b200.ml import tensorcore as TC
    pair_mode(pair)
    g = torch.Generator(device="cuda").manual_seed(m * n + k)
    a = torch.randn((k, m) if ta else (m, k), device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn((n, k) if tb else (k, n), device="cuda", generator=g).to(torch.bfloat16)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    scale = ref.abs().max().item()
    c = torch.full((m, n), 7.0, device="cuda")
    TC.gemm(a, b, c, transa=bool(ta), transb=bool(tb), precision="bf16", split_k=1, tile_n=256 if n > 128 else 128)
    assert (c.double() - ref).abs().max().item() / scale < 2e-5 + 4e-9 * k
    bias = torch.randn(n, device="cuda", generator=g)
    c2 = torch.full((m, n), 7.0, device="cuda")
    TC.gemm(a, b, c2, transa=bool(ta), transb=bool(tb), bias=bias, relu=True, precision="bf16")
    assert (c2.double() - torch.relu(ref + bias.double())).abs().max().item() / scale < 2e-5 + 4e-9 * k


def test_bf16_mode_pieces():
    """The extra pieces of the "bf16" mode: fp32 -> bf16 conversion (round to nearest even), the fused BCE epilogue
    writing bf16 dlogits (+ bias gradient from the rounded values), the noise kernel's bf16 dense rows, the
    persistent softmax-KL kernel writing bf16 dlogits."""
    from cubecobrarecommender_b200.ml import tensorcore as TC
    g = torch.Generator(device="cuda").manual_seed(8)
    # (1) conversion
    src = torch.randn(37, 520, device="cuda", generator=g)
    dst = torch.zeros(37, 512, dtype=torch.bfloat16, device="cuda")
    E.call("cc_convert_f32_bf16", E.ptr(src), src.stride(0), E.ptr(dst), dst.stride(0), 37, 512, E.stream_ptr())
    assert torch.equal(dst, src[:, :512].to(torch.bfloat16))
    # (2) fused BCE with bf16 dlogits
    m, k, c = 200, 512, 1000
    cpad = 1024
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(k, cpad, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    bias = torch.randn(c, device="cuda", generator=g)
    y = (torch.rand(m, c, device="cuda", generator=g) < 0.05).double().cpu().numpy()
    yb = torch.tensor(_bits_from_dense(y)).cuda()
    dz = torch.full((m, cpad), 9.0, dtype=torch.bfloat16, device="cuda")
    part = torch.zeros(TC.bce_partial_count(m, cpad), dtype=torch.float64, device="cuda")
    dbias = torch.full((c,), 5.0, device="cuda")
    TC.gemm_bce(a, w[:, :c], bias, yb, float(m * c), dz, part, precision="bf16", dbias=dbias)
    z = (a.double() @ w[:, :c].double() + bias.double()).cpu().numpy()
    ref_loss = od.bce_from_logits_np(z, y)
    assert abs(part.sum().item() / (m * c) - ref_loss) / ref_loss < 1e-5
    ref_dz = (1 / (1 + np.exp(-z)) - y) / (m * c)
    got = dz[:, :c].float().cpu().numpy()
    assert np.abs(got - ref_dz).max() <= 2.0 ** -8 * np.abs(ref_dz).max()          # bf16: 8 significant bits
    assert (dz[:, c:] == 0).all()
    assert np.abs(dbias.cpu().numpy() - got.astype(np.float64).sum(0)).max() < 1e-5 * np.abs(got.sum(0)).max() + 1e-12
    # (3) noise kernel: bf16 dense rows == float dense rows
    from cubecobrarecommender_b200 import graph as G
    cc, kk, b = 384, 96, 32
    ip, ix = synth_cubes_csr(kk, cc, size_lo=12, size_hi=70, seed=7)
    csr = CubeCSR(ip, ix, cc)
    gr = G.build_graph(csr, "cuda", want_m64=False)
    prob, alias = E.alias_table(gr.neg_sampler.cpu().numpy(), "cuda")
    indptr, indices = G.upload_csr(csr, "cuda")
    dense = {}
    for prec in ("tf32", "bf16"):
        model = M.CC_Recommender(cc, device="cuda", precision=prec)
        eng = E.DAEEngine(model, gr.mhat, batch=b, reg_rows=b, reg=0.1, max_cube_size=80)
        eng.sample_batch(indptr, indices, torch.arange(b, dtype=torch.int32, device="cuda"), prob, alias, seed=5)
        dense[prec] = eng.x_dense.float().clone()
        if prec == "bf16":
            loss = eng.train_step().cpu().numpy()          # one whole step runs (bf16 dlogits out of the softmax-KL kernel)
            assert np.isfinite(loss).all()
    assert torch.equal(dense["tf32"], dense["bf16"]) and dense["bf16"].sum() > 0


@pytest.mark.parametrize("c,rows", [(340, 700), (20884, 333), (22528, 160)])
@pytest.mark.parametrize("fast", [0, 1])
def test_softmax_kl_persistent_kernels_vs_oracle(c, rows, fast):
    """The two persistent softmax-KL kernels (512 threads with the target row and the column sums in registers -- the
    shipped one up to 22 528 cards -- and the 1024-thread form) against the float64 oracle and against each other, with
    and without the per-target-row table, float32 and bf16 dlogits, more rows than CTAs, rows with clipped entries."""
    from cubecobrarecommender_b200 import _lib
    cpad = (c + 127) // 128 * 128
    assert _lib.load().cc_softmax_kl_fuses_dbias(c, cpad, cpad, c, cpad)
    rng = np.random.default_rng(c + rows)
    g = torch.Generator(device="cuda").manual_seed(c)
    z = torch.randn(rows, cpad, device="cuda", generator=g) * 3
    z[1, 5] = 45.0                                           # drives most of row 1 below the 1e-7 clip
    z[:, c:] = 0
    nt = 64                                                  # distinct target rows
    t = torch.rand(nt, c, device="cuda", generator=g)
    t = torch.where(t < 0.7, torch.zeros_like(t), t)
    t = (t / t.sum(1, keepdim=True)).contiguous()
    tr = torch.from_numpy(rng.integers(0, nt, size=rows).astype(np.int32)).cuda()
    table = torch.zeros(nt, dtype=torch.float64, device="cuda")
    E.call("cc_kl_target_table", E.ptr(t), c, nt, c, E.ptr(table), E.stream_ptr())
    scale = 0.1 / rows
    # oracle on a sample of rows (float64)
    pick = np.unique(np.concatenate([[0, 1, rows - 1], rng.integers(0, rows, size=12)]))
    zz = z[torch.from_numpy(pick).cuda(), :c].double().cpu().numpy()
    tt = t[tr[torch.from_numpy(pick).cuda()].long()].double().cpu().numpy()
    q = od.softmax_np(zz)
    tcl = np.clip(tt, 1e-7, 1.0); qcl = np.clip(q, 1e-7, 1.0)
    ref_rows = (tcl * np.log(tcl / qcl)).sum(1)
    un = (q >= 1e-7) & (q <= 1)
    ref_dz = scale * (q * (tcl * un).sum(1, keepdims=True) - tcl * un)
    outs = {}
    try:
        for variant in (0, 1):
            E.call("cc_softmax_kl_set_variant", variant)
            for use_table in (False, True):
                dz = torch.full((rows, cpad), 9.0, device="cuda")
                rl = torch.zeros(rows, dtype=torch.float64, device="cuda")
                db = torch.full((c,), 3.0, device="cuda")
                E.call("cc_softmax_kl_fwd_bwd_ex", E.ptr(z), cpad, E.ptr(t), c, E.ptr(tr), rows, c, cpad, scale, E.ptr(dz), cpad,
                       E.ptr(rl), fast, E.ptr(db), None, 0, E.ptr(table) if use_table else None, E.stream_ptr())
                outs[(variant, use_table)] = (dz, rl, db)
                got_rows = rl[torch.from_numpy(pick).cuda()].cpu().numpy()
                assert np.abs(got_rows - ref_rows).max() < (2e-5 if fast else 5e-6) * np.abs(ref_rows).max()
                got_dz = dz[torch.from_numpy(pick).cuda(), :c].double().cpu().numpy()
                edge = np.abs(q - 1e-7) < 1e-10
                tol = (2e-3 if fast else 2e-5) * np.abs(ref_dz).max()      # fast: MUFU exp + tf32-rounded dlogits
                assert np.abs(got_dz - ref_dz)[~edge].max() < tol
                assert (dz[:, c:] == 0).all()
                ref_db = dz[:, :c].double().sum(0)
                assert (db.double() - ref_db).abs().max().item() < 2e-6 * ref_db.abs().max().item() + 1e-12
            # bf16 dlogits
            dz16 = torch.full((rows, cpad), 9.0, dtype=torch.bfloat16, device="cuda")
            rl16 = torch.zeros(rows, dtype=torch.float64, device="cuda")
            db16 = torch.zeros(c, device="cuda")
            E.call("cc_softmax_kl_fwd_bwd_ex", E.ptr(z), cpad, E.ptr(t), c, E.ptr(tr), rows, c, cpad, scale, None, 0,
                   E.ptr(rl16), 1, E.ptr(db16), E.ptr(dz16), cpad, E.ptr(table), E.stream_ptr())
            ref16 = outs[(variant, True)][0] if fast else None
            assert (dz16[:, c:] == 0).all()
            if fast:        # the same float values before the bf16 rounding, up to the skipped tf32 rounding
                assert (dz16.float() - ref16).abs().max().item() <= 2.0 ** -8 * ref16.abs().max().item()
        # the two kernels: same formulas, different reduction trees
        for use_table in (False, True):
            a, b = outs[(0, use_table)], outs[(1, use_table)]
            assert torch.allclose(a[0], b[0], rtol=3e-3 if fast else 3e-5, atol=1e-12)
            assert torch.allclose(a[1], b[1], rtol=1e-5)
    finally:
        E.call("cc_softmax_kl_set_variant", 0)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_keras_accuracy_metrics_vs_oracle(precision):
    """metrics=['accuracy'] (reference train.py:87): binary accuracy of the sigmoid tower (counted in the fused BCE
    epilogue, or by cc_binary_accuracy_rows in the exact-fp32 mode) and categorical accuracy of the softmax tower (first
    maximal logit against the target row's argmax, counted in the softmax-KL kernel), against the float64 oracle."""
    c, x, y, rows, mh = _problem(c=400, k=300, b=96, r=80, seed=21)
    params = od.init_params(c, seed=3)
    rng = np.random.default_rng(2)
    for kname in params:          # larger weights: logits away from zero, peaked softmax rows
        params[kname] = (params[kname] * 3 + (rng.standard_normal(params[kname].shape) * 0.2 if kname.endswith("bias") else 0)).astype(np.float32)
    model = M.CC_Recommender(c, device="cuda", precision=precision)
    model.set_weights_dict(params)
    mhat = torch.tensor(mh.astype(np.float32)).cuda()
    eng = E.DAEEngine(model, mhat, batch=x.shape[0], reg_rows=len(rows), reg=0.1, max_cube_size=80, metrics=True)
    sb = M.SparseBatch.from_csr(CubeCSR.from_dense(x), "cuda")
    eng.set_batch(sb, torch.tensor(_bits_from_dense(y)).cuda(), torch.tensor(rows).cuda())
    eng.forward_backward()
    got = eng.metrics2.cpu().numpy()
    z1, z2, _ = od.forward_np(params, x, rows)
    t32 = mh.astype(np.float32)[rows]
    ref1, ref2 = od.binary_accuracy_np(z1, y), od.categorical_accuracy_np(z2, t32)
    # cells whose logit is within the mode's rounding of zero, rows whose two best logits are that close, may flip
    band = {"fp32": 1e-5, "tf32": 5e-3, "bf16": 3e-2}[precision]
    near1 = float(np.mean(np.abs(z1) < band * np.abs(z1).max()))
    top2 = np.sort(z2, axis=1)[:, -2:]
    near2 = float(np.mean((top2[:, 1] - top2[:, 0]) < band * np.abs(z2).max()))
    assert abs(got[0] - ref1) <= near1 + 1e-12, (got, ref1, near1)
    assert abs(got[1] - ref2) <= near2 + 1e-12, (got, ref2, near2)
    assert 0.5 < ref1 < 1.0
    # the table behind the target side: first maximal column of every M-hat row
    assert np.array_equal(eng.kl_argmax().cpu().numpy(), np.argmax(mh.astype(np.float32), axis=1))
    # metrics ride along without touching losses or gradients
    ref_eng = E.DAEEngine(M.CC_Recommender(c, device="cuda", precision=precision), mhat, batch=x.shape[0], reg_rows=len(rows),
                          reg=0.1, max_cube_size=80)
    ref_eng.model.set_weights_dict(params)
    ref_eng.set_batch(sb, torch.tensor(_bits_from_dense(y)).cuda(), torch.tensor(rows).cuda())
    ref_eng.forward_backward()
    assert torch.allclose(ref_eng.loss3, eng.loss3, rtol=1e-12)
