"""CUDA graph build / scoring / top-N vs the oracle and the golden fixtures, through the C ABI.
Bar: counts and ids bit-exact, M bit-exact in float64 (<=1e-6 rel required), M-hat <=1e-6 rel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from cubecobrarecommender_b200 import graph as G
from cubecobrarecommender_b200.sparse import CubeCSR
from cubecobrarecommender_b200.synth import csr_to_dense, synth_cubes_csr
from oracle import graph as og


def _golden_csr(g):
    return CubeCSR(g["indptr"], g["indices"], int(g["num_cards"]))


def test_golden_adjacency_host_entry(graph_golden):
    g = graph_golden
    csr = _golden_csr(g)
    m, counts = G.create_adjacency_matrix_host(csr, return_counts=True)
    assert np.array_equal(m, g["adj"])                                  # bit-exact float64
    assert np.array_equal(counts, og.cooc_counts(csr.indptr, csr.indices, csr.num_cards))
    m2 = G.create_adjacency_matrix_host(csr, force_diag=0.5)
    assert np.array_equal(m2, g["adj_force_diag"])


METHODS = ["tensor", "popcount"]     # tcgen05 kind::i8 contraction | bit-packed AND+POPC tiles


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("k,c,lo,hi", [(1, 5, 1, 5), (33, 130, 3, 60), (700, 1000, 20, 200), (2100, 300, 5, 40),
                                       (300, 515, 100, 400)])
def test_counts_and_normalise_vs_oracle(k, c, lo, hi, method):
    ip, ix = synth_cubes_csr(k, c, size_lo=lo, size_hi=hi, seed=k + c)
    csr = CubeCSR(ip, ix, c)
    gr = G.build_graph(csr, "cuda", method=method)
    cnt = og.cooc_counts(ip, ix, c)
    assert np.array_equal(gr.counts.cpu().numpy(), cnt)
    m = og.adjacency_from_counts(cnt)
    assert np.array_equal(gr.m64.cpu().numpy(), m)
    mh = og.m_hat(m)
    got = gr.mhat.cpu().numpy().astype(np.float64)
    assert np.abs(got - mh).max() <= 1e-6 * mh.max()
    nz = mh > 0
    assert (np.abs(got[nz] - mh[nz]) / mh[nz]).max() < 1e-6
    ns = og.neg_sampler(mh)
    assert np.abs(gr.neg_sampler.cpu().numpy() - ns).max() < 1e-12
    assert abs(gr.neg_sampler.sum().item() - 1) < 1e-12


@pytest.mark.parametrize("method", METHODS)
def test_empty_and_ragged_cubes(method):
    # empty cubes, a cube with every card, duplicate ids collapse
    csr = CubeCSR.from_lists([[], [0, 1, 2, 3, 4, 5, 6], [3, 3, 3], [], [6, 0]], 7)
    gr = G.build_graph(csr, "cuda", method=method)
    cnt = og.cooc_counts(csr.indptr, csr.indices, 7)
    assert np.array_equal(gr.counts.cpu().numpy(), cnt)
    with pytest.raises(ValueError):
        bad = CubeCSR(np.array([0, 2]), np.array([1, 9], dtype=np.int32), 7)
        G.build_graph(bad, "cuda", method=method)


@pytest.mark.parametrize("method", METHODS)
def test_accumulate_equals_single_pass(method):
    ip, ix = synth_cubes_csr(300, 257, size_lo=5, size_hi=50, seed=5)
    csr = CubeCSR(ip, ix, 257)
    whole = G.build_graph(csr, "cuda", want_m64=False, want_mhat=False, want_neg=False, method="popcount").counts
    acc = None
    for r in range(3):
        sh = csr.shard(r, 3)
        a, b = G.upload_csr(sh, "cuda")
        acc = G.count_cooccurrence(a, b, sh.num_cubes, 257, counts=acc, accumulate=acc is not None, method=method)
    assert torch.equal(acc, whole)          # "checksum of checksums": shards sum to the whole


def test_tensor_count_chunks_over_cubes():
    """More cubes than one pass of the byte matrix holds (cc_cooc_tc_chunk_cubes): passes accumulate."""
    from cubecobrarecommender_b200 import _lib
    k, c = 40000, 264
    assert _lib.load().cc_cooc_tc_chunk_cubes(k) < k
    ip, ix = synth_cubes_csr(k, c, size_lo=2, size_hi=30, seed=9)
    a, b = G.upload_csr(CubeCSR(ip, ix, c), "cuda")
    t = G.count_cooccurrence(a, b, k, c, method="tensor")
    p = G.count_cooccurrence(a, b, k, c, method="popcount")
    assert torch.equal(t, p)
    assert np.array_equal(t.cpu().numpy(), og.cooc_counts(ip, ix, c))


def test_size_independent_properties_large():
    ip, ix = synth_cubes_csr(4096, 4000, cfg=1)
    csr = CubeCSR(ip, ix, 4000)
    gr = G.build_graph(csr, "cuda", method="tensor")
    cnt = gr.counts
    assert torch.equal(cnt, G.build_graph(csr, "cuda", method="popcount", want_m64=False, want_mhat=False,
                                          want_neg=False).counts)      # the two count kernels agree bit for bit
    assert torch.equal(cnt, cnt.t())                                       # symmetric
    sizes = torch.from_numpy(np.diff(ip)).cuda()
    assert int(cnt.diagonal().sum()) == int(sizes.sum())                   # diag = card frequency
    assert int(cnt.sum()) == int((sizes * sizes).sum())                    # sum = sum of s^2
    d = gr.m64.diagonal()
    assert set(d.unique().tolist()) <= {0.0, 1.0}
    assert torch.allclose(gr.mhat.double().sum(1), torch.ones(4000, dtype=torch.float64, device="cuda"), atol=1e-5)


def _rank_check(scores, ids, expected_ids):
    ids = np.asarray(ids); expected_ids = np.asarray(expected_ids)
    assert np.array_equal(ids, expected_ids)


def test_golden_recs_and_cuts(graph_golden, pairwise_golden):
    for g, live in ((graph_golden, 120), (pairwise_golden, None)):
        c = int(g["num_cards"])
        csr_all = CubeCSR(g["indptr"], g["indices"], c)
        dense = csr_to_dense(g["indptr"], g["indices"], c) if live is None else \
            np.pad(csr_to_dense(g["indptr"], g["indices"], live), ((0, 0), (0, c - live)))
        adj = og.create_adjacency_matrix(dense)
        rec = G.GraphRecommender(torch.from_numpy(adj).cuda())
        rows = g["rec_cube_rows"]
        sub = csr_all.rows(rows)
        scores, _, _ = rec.scores(sub)
        scores = scores.cpu().numpy()
        for n, r in enumerate(rows):
            missing = dense[r] == 0
            assert np.array_equal(scores[n][missing], g["rec_scores"][n][missing])   # bit-identical float64
        ids, vals, cnt = rec.recs(sub, 50)
        ids = ids.cpu().numpy()
        for n, r in enumerate(rows):
            expect = np.array(og.simple_recs(dense[r], adj))[:50]
            assert np.array_equal(ids[n], expect)
            # the reference's own (unstable-sort) answer: same scores in the same order
            s = g["rec_scores"][n]
            assert np.array_equal(s[ids[n]], s[g["recs_top50"][n]])
        ncut = int(dense[rows].sum(1).max())
        cids, cvals, ccnt = rec.cuts(sub, ncut)
        cids = cids.cpu().numpy(); ccnt = ccnt.cpu().numpy()
        for n, r in enumerate(rows):
            expect = np.array(og.simple_cuts(dense[r], adj.copy()))
            assert ccnt[n] == len(expect)
            assert np.array_equal(cids[n][:ccnt[n]], expect)
            assert (cids[n][ccnt[n]:] == -1).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 50, 128, 129, 2048, 3000])
def test_topn_masked_ties_and_full_ranking(dtype, n):
    """float32 with n <= 128 runs the warp-per-cube streaming select, larger n / float64 the radix select and the
    full bitonic ranking: all three against numpy's stable argsort under the documented tie rule."""
    rng = np.random.default_rng(7)
    c, batch = 5000, 6
    # heavy ties: scores drawn from 40 distinct values, plus -0.0/+0.0
    vals = rng.integers(0, 40, size=(batch, c)).astype(np.float64) / 8 - 2
    vals[0, :100] = -0.0
    scores = torch.tensor(vals, dtype=dtype).cuda()
    lists = [np.sort(rng.choice(c, size=rng.integers(0, 600), replace=False)) for _ in range(batch)]
    csr = CubeCSR.from_lists(lists, c)
    mp = torch.from_numpy(csr.indptr).cuda(); mi = torch.from_numpy(csr.indices).cuda()
    if len(csr.indices) == 0:
        mi = torch.zeros(1, dtype=torch.int32).cuda()
    sv = scores.cpu().numpy()
    ids, v, cnt = G.topn_masked(scores, mp, mi, n, only_listed=False, descending=True)
    ids = ids.cpu().numpy(); cnt = cnt.cpu().numpy(); v = v.cpu().numpy()
    for b in range(batch):
        order = sv[b].argsort(kind="stable")[::-1]
        inm = np.zeros(c, bool); inm[lists[b]] = True
        expect = [i for i in order if not inm[i]][:n]
        assert cnt[b] == len(expect)
        assert np.array_equal(ids[b][:cnt[b]], expect)
        assert np.array_equal(v[b][:cnt[b]], sv[b][expect])
    ids, v, cnt = G.topn_masked(scores, mp, mi, n, only_listed=True, descending=False)
    ids = ids.cpu().numpy(); cnt = cnt.cpu().numpy()
    for b in range(batch):
        order = sv[b].argsort(kind="stable")
        inm = np.zeros(c, bool); inm[lists[b]] = True
        expect = [i for i in order if inm[i]][:n]
        assert cnt[b] == len(expect)
        assert np.array_equal(ids[b][:cnt[b]], expect)
        assert (ids[b][cnt[b]:] == -1).all()


@pytest.mark.parametrize("c,batch,n", [(5000, 6, 50), (4999, 11, 128), (97, 3, 7), (20884, 9, 50)])
def test_topn_streaming_select_equals_radix_select(c, batch, n):
    """The two float32 kernels implement one total order (score, then index): identical ids and values on
    heavy ties, on ascending-sorted rows (every element beats the running threshold: worst case for the
    streaming select), on descending rows, and when the mask leaves fewer than n candidates."""
    from cubecobrarecommender_b200 import _lib
    rng = np.random.default_rng(c + n)
    vals = rng.integers(0, 30, size=(batch, c)).astype(np.float32) / 4 - 3
    vals[0] = np.sort(rng.standard_normal(c).astype(np.float32))              # ascending
    if batch > 1:
        vals[1] = np.sort(rng.standard_normal(c).astype(np.float32))[::-1]    # descending
    lists = [np.sort(rng.choice(c, size=rng.integers(0, min(c, 700)), replace=False)) for _ in range(batch)]
    lists[-1] = np.arange(3, c)                                               # only 3 candidates left
    csr = CubeCSR.from_lists(lists, c)
    mp = torch.from_numpy(csr.indptr).cuda(); mi = torch.from_numpy(csr.indices).cuda()
    scores = torch.from_numpy(vals).cuda()
    for only_listed, desc in ((False, True), (True, False), (False, False), (True, True)):
        a = G.topn_masked(scores, mp, mi, n, only_listed=only_listed, descending=desc)
        _lib.call("cc_topn_set_force_radix", 1)
        try:
            b = G.topn_masked(scores, mp, mi, n, only_listed=only_listed, descending=desc)
        finally:
            _lib.call("cc_topn_set_force_radix", 0)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # numpy cross-check of the descending / not-listed case
    ids, v, cnt = (t.cpu().numpy() for t in G.topn_masked(scores, mp, mi, n))
    for r in range(batch):
        order = vals[r].argsort(kind="stable")[::-1]
        inm = np.zeros(c, bool); inm[lists[r]] = True
        expect = [i for i in order if not inm[i]][:n]
        assert cnt[r] == len(expect) and np.array_equal(ids[r][:cnt[r]], expect)
        assert (ids[r][cnt[r]:] == -1).all()


def test_topn_fused_sigmoid_equals_sigmoid_then_select():
    """cc_topn_masked_sigmoid_f32 ranks float32 sigmoid(logit) computed on the fly: same ids and the same
    probability bits as cc_sigmoid_f32 followed by cc_topn_masked_f32 (saturated logits tie at exactly 1.0)."""
    from cubecobrarecommender_b200._lib import call, ptr, stream_ptr
    rng = np.random.default_rng(3)
    c, batch, n = 3001, 13, 50
    z = (rng.standard_normal((batch, c)) * 12).astype(np.float32)            # many saturate to 1.0 / underflow to 0
    lists = [np.sort(rng.choice(c, size=rng.integers(1, 500), replace=False)) for _ in range(batch)]
    csr = CubeCSR.from_lists(lists, c)
    mp = torch.from_numpy(csr.indptr).cuda(); mi = torch.from_numpy(csr.indices).cuda()
    logits = torch.from_numpy(z).cuda()
    fused = G.topn_masked(logits, mp, mi, n, sigmoid=True)
    probs = torch.empty_like(logits)
    call("cc_sigmoid_f32", ptr(logits), ptr(probs), logits.numel(), stream_ptr())
    two_pass = G.topn_masked(probs, mp, mi, n)
    for x, y in zip(fused, two_pass):
        assert torch.equal(x, y)
    assert (fused[1] == 1.0).any()          # the tie rule was exercised on saturated scores
    # every mode of the select (the logit pre-filter has a lower-bound and an upper-bound form), rows sorted both
    # ways (worst / best case for the running threshold) and a row of identical logits
    z2 = z.copy()
    z2[0] = np.sort(z2[0]); z2[1] = np.sort(z2[1])[::-1]; z2[2] = 0.25
    logits2 = torch.from_numpy(z2).cuda()
    call("cc_sigmoid_f32", ptr(logits2), ptr(probs), logits2.numel(), stream_ptr())
    for only_listed, desc in ((False, True), (True, False), (False, False), (True, True)):
        a = G.topn_masked(logits2, mp, mi, n, sigmoid=True, only_listed=only_listed, descending=desc)
        b = G.topn_masked(probs, mp, mi, n, only_listed=only_listed, descending=desc)
        for x, y in zip(a, b):
            assert torch.equal(x, y), (only_listed, desc)
    # an untrained model: every logit of a row within a few hundredths of the others (and of zero, of +-9, of 15, where
    # the sigmoid is steep / flat) -- the regime in which the float32 logit bound of the row select is tightest
    for centre, spread in ((0.0, 0.03), (-9.0, 0.02), (9.0, 0.05), (15.5, 0.5), (0.0, 1e-4)):
        z3 = (centre + rng.standard_normal((batch, c)) * spread).astype(np.float32)
        logits3 = torch.from_numpy(z3).cuda()
        call("cc_sigmoid_f32", ptr(logits3), ptr(probs), logits3.numel(), stream_ptr())
        padded = torch.zeros((batch, c + 3), device="cuda")      # row stride % 4 == 0: the CTA-per-cube row select
        padded[:, :c] = logits3
        for desc in (True, False):
            b = G.topn_masked(probs, mp, mi, n, descending=desc)
            for src in (logits3, padded[:, :c]):                 # (stride 3001: the warp-per-cube streaming select)
                a = G.topn_masked(src, mp, mi, n, sigmoid=True, descending=desc)
                for x, y in zip(a, b):
                    assert torch.equal(x, y), (centre, spread, desc, src.stride(0))


def _algo(a):
    from cubecobrarecommender_b200 import _lib
    _lib.call("cc_topn_set_algo", a)


# the forms of the row select that are compared with the streaming select: 2 / 3 = both sweeps over shared memory (one
# CTA per SM with two row buffers / two CTAs per SM), 4 = the register form (the default when the row fits)
ROW_SELECT_ALGOS = (2, 3, 4)


@pytest.mark.parametrize("c,ld,batch,n", [(5000, 5000, 6, 50), (4999, 5008, 11, 128), (97, 100, 3, 7),
                                          (20884, 20992, 9, 50), (2001, 2004, 700, 50), (20884, 20992, 5, 1)])
def test_topn_row_select_equals_streaming_select(c, ld, batch, n):
    """The CTA-per-cube row select (rows staged in shared memory by bulk copies, two sweeps) against the warp-per-cube
    streaming select and numpy's stable argsort: heavy ties, rows sorted both ways, a row whose large values all sit
    in 40 of the 512 thread strides (more than RS_CAP survivors: the threshold is raised and the sweep repeated),
    masks that leave fewer than n candidates, duplicate and out-of-range list entries, padded row strides whose
    padding holds huge values, and more cubes than two per CTA (both row buffers and both barrier phases reused)."""
    rng = np.random.default_rng(c + n + batch)
    vals = rng.integers(0, 30, size=(batch, c)).astype(np.float32) / 4 - 3
    vals[0] = np.sort(rng.standard_normal(c).astype(np.float32))
    vals[1] = np.sort(rng.standard_normal(c).astype(np.float32))[::-1]
    vals[2] = rng.standard_normal(c).astype(np.float32)
    vals[2][(np.arange(c) % 512) < 40] += 100.0
    if batch > 4:
        vals[3] = rng.standard_normal(c).astype(np.float32)
        vals[4] = -vals[2]
        vals[3][rng.random(c) < 0.7] = -np.inf                                # most leaders are infinities
        vals[3][rng.random(c) < 0.05] = -0.0
        vals[3][rng.random(c) < 0.05] = 0.0
    if batch > 100:
        vals[5:] = rng.standard_normal((batch - 5, c)).astype(np.float32)
        vals[6][rng.random(c) < 0.7] = np.inf
    lists = [np.sort(rng.choice(c, size=rng.integers(0, min(c, 700)), replace=False)) for _ in range(batch)]
    lists[-1] = np.arange(3, c)                                               # only 3 candidates left
    lists[0] = np.concatenate([lists[0], lists[0][:5], [-1, c, c + 7]]).astype(np.int64)   # duplicates, out of range
    ip = np.zeros(batch + 1, np.int64); ip[1:] = np.cumsum([len(x) for x in lists])
    ix = np.concatenate(lists).astype(np.int32)
    mp = torch.from_numpy(ip).cuda(); mi = torch.from_numpy(ix).cuda()
    full = torch.full((batch, ld), 1e30, dtype=torch.float32, device="cuda")
    full[:, :c] = torch.from_numpy(vals).cuda()
    scores = full[:, :c]
    assert scores.stride(0) == ld
    try:
        for only_listed, desc in ((False, True), (True, False), (False, False), (True, True)):
            _algo(1)
            b = G.topn_masked(scores, mp, mi, n, only_listed=only_listed, descending=desc)
            for variant in ROW_SELECT_ALGOS:
                _algo(variant)
                a = G.topn_masked(scores, mp, mi, n, only_listed=only_listed, descending=desc)
                for x, y in zip(a, b):
                    assert torch.equal(x, y), (only_listed, desc, variant)
        _algo(2)
        ids, v, cnt = (t.cpu().numpy() for t in G.topn_masked(scores, mp, mi, n))
        # fused sigmoid: the same rows as logits
        logits = full.clone(); logits[:, :c] *= 6.0
        _algo(1)
        b = G.topn_masked(logits[:, :c], mp, mi, n, sigmoid=True)
        for variant in ROW_SELECT_ALGOS:
            _algo(variant)
            a = G.topn_masked(logits[:, :c], mp, mi, n, sigmoid=True)
            for x, y in zip(a, b):
                assert torch.equal(x, y), variant
    finally:
        _algo(0)
    for r in range(min(batch, 12)):
        order = vals[r].argsort(kind="stable")[::-1]
        inm = np.zeros(c, bool); inm[[i for i in lists[r] if 0 <= i < c]] = True
        expect = [i for i in order if not inm[i]][:n]
        assert cnt[r] == len(expect) and np.array_equal(ids[r][:cnt[r]], expect)
        assert np.array_equal(v[r][:cnt[r]], vals[r][expect])
        assert (ids[r][cnt[r]:] == -1).all()


def test_topn_row_select_is_the_default_for_padded_rows():
    """Rows that qualify (ld % 4 == 0, aligned) take the row select automatically; algo 2 on rows that do not is an error."""
    from cubecobrarecommender_b200 import _lib
    rng = np.random.default_rng(1)
    scores = torch.from_numpy(rng.standard_normal((4, 999)).astype(np.float32)).cuda()      # ld = 999: not eligible
    mp = torch.zeros(5, dtype=torch.int64, device="cuda"); mi = torch.zeros(1, dtype=torch.int32, device="cuda")
    ids, _, _ = G.topn_masked(scores, mp, mi, 10)                       # automatic: streaming select
    assert np.array_equal(ids.cpu().numpy(), np.argsort(-scores.cpu().numpy(), axis=1, kind="stable")[:, :10])
    _algo(2)
    try:
        with pytest.raises(_lib.CubeCobraError):
            G.topn_masked(scores, mp, mi, 10)
    finally:
        _algo(0)


def test_scale_up_card_count_properties():
    """configs[4] card count (C = 100 000; the reference cannot even allocate this dense): one pass of the tensor-core
    count kernel over 8 192 cubes, checked through size-independent properties (the oracle would take hours):
    diag = card frequency, total = sum of s^2, symmetry on a million sampled pairs, and agreement with the popcount
    kernel on a 2 048 x 2 048 corner block."""
    k, c = 8192, 100_000
    ip, ix = synth_cubes_csr(k, c, size_lo=360, size_hi=720, seed=31)
    a, b = G.upload_csr(CubeCSR(ip, ix, c), "cuda")
    cnt = G.count_cooccurrence(a, b, k, c, method="tensor")
    sizes = torch.from_numpy(np.diff(ip)).cuda()
    freq = torch.bincount(b.long(), minlength=c)
    assert torch.equal(cnt.diagonal().long(), freq)
    assert int(cnt.sum(dtype=torch.int64)) == int((sizes.long() ** 2).sum())
    g = torch.Generator(device="cuda").manual_seed(1)
    i = torch.randint(0, c, (1_000_000,), device="cuda", generator=g)
    j = torch.randint(0, c, (1_000_000,), device="cuda", generator=g)
    assert torch.equal(cnt[i, j], cnt[j, i])
    # popular cards (low ids under the Zipf law) against the popcount kernel restricted to those cards
    keep = b < 2048
    sub_len = torch.zeros(k, dtype=torch.int64, device="cuda").scatter_add_(
        0, torch.repeat_interleave(torch.arange(k, device="cuda"), sizes.long())[keep], torch.ones(int(keep.sum()), dtype=torch.int64, device="cuda"))
    sub_ip = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), sub_len.cumsum(0)])
    ref = G.count_cooccurrence(sub_ip, b[keep].contiguous(), k, 2048, method="popcount")
    assert torch.equal(cnt[:2048, :2048], ref)


@pytest.mark.parametrize("c,world", [(515, 4), (1000, 8), (64, 3)])
def test_row_block_normalise_equals_whole_matrix(c, world):
    """The reduce-scattered form of the sharded build (configs[4]): every rank normalises only its row block of the summed
    counts (cc_row_normalise_rows / cc_col_mass_rows / cc_col_mass_scale).  Emulated on one GPU: the blocks'
    M / M-hat / row sums are bit-identical to the whole-matrix pass, the summed column masses give the same sampler."""
    from cubecobrarecommender_b200._lib import call, ptr, stream_ptr
    ip, ix = synth_cubes_csr(300, c, size_lo=3, size_hi=min(c, 90), seed=c + world)
    csr = CubeCSR(ip, ix, c)
    indptr, indices = G.upload_csr(csr, "cuda")
    counts = G.alloc_counts(c, world)
    G.count_cooccurrence(indptr, indices, csr.num_cubes, c, counts=counts)
    whole = G.normalise(counts, want_m64=True, want_mhat=True, want_neg=True)
    assert np.array_equal(whole.counts.cpu().numpy(), og.cooc_counts(ip, ix, c))
    blk = -(-c // world)
    mass = torch.zeros(c, dtype=torch.float64, device="cuda")
    for r in range(world):
        r0 = r * blk
        nrows = max(0, min(blk, c - r0))
        part = G.normalise_rows(counts[r0:r0 + nrows], r0, c, want_m64=True)          # (no process group: no all_reduce)
        assert torch.equal(part.m64, whole.m64[r0:r0 + nrows])
        assert torch.equal(part.mhat, whole.mhat[r0:r0 + nrows, :c])
        assert torch.equal(part.rowsum, whole.rowsum[r0:r0 + nrows])
        # normalise_rows scales its own block's masses; undo that to emulate the all_reduce of the raw masses
        ws = torch.empty(G._lib.load().cc_col_mass_workspace_bytes(c) // 8, dtype=torch.float64, device="cuda")
        raw = torch.empty(c, dtype=torch.float64, device="cuda")
        call("cc_col_mass_rows", ptr(counts[r0:r0 + nrows]), counts.stride(0), r0, nrows, c, ptr(part.rowsum), ptr(ws), ptr(raw),
             stream_ptr())
        mass += raw
    call("cc_col_mass_scale", ptr(mass), c, stream_ptr())
    assert (mass - whole.neg_sampler).abs().max().item() < 1e-14
