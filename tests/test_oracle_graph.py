"""Oracle vs the reference's own outputs (tests/golden/graph_small.npz) and
hand-computed known answers for reference src/non_ml/utils.py:75-92,
src/ml/train.py:69-71, src/scripts/recommend.py:7-18, cut_cards.py:7-18."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from cubecobrarecommender_b200.synth import csr_to_dense, dense_to_csr, synth_cubes_csr
from oracle import graph


def test_golden_adjacency_bit_exact(graph_golden):
    g = graph_golden
    c = int(g["num_cards"])
    dense = np.zeros((len(g["indptr"]) - 1, c))
    dense[:, :120] = csr_to_dense(g["indptr"], g["indices"], 120)
    assert np.array_equal(graph.create_adjacency_matrix(dense), g["adj"])
    assert np.array_equal(graph.create_adjacency_matrix_loop(dense), g["adj"])
    assert np.array_equal(graph.create_adjacency_matrix(dense, force_diag=0.5), g["adj_force_diag"])
    # zero-count cards -> all-zero rows (utils.py:87-88)
    assert not g["adj"][120:].any()


def test_golden_recs_and_cuts(graph_golden):
    g = graph_golden
    c = int(g["num_cards"])
    dense = np.zeros((len(g["indptr"]) - 1, c))
    dense[:, :120] = csr_to_dense(g["indptr"], g["indices"], 120)
    adj = g["adj"]
    for n, row in enumerate(g["rec_cube_rows"]):
        cube = dense[row]
        scores = graph.simple_recs_scores(cube, adj)
        missing = np.where(cube == 0)[0]
        assert np.array_equal(scores[missing], g["rec_scores"][n][missing])
        ours = np.array(graph.simple_recs(cube, adj))[:50]
        ref = g["recs_top50"][n]
        # identical score sequence; identical ids wherever scores are distinct
        assert np.array_equal(scores[ours], scores[ref])
        distinct = np.ones(50, bool)
        s = scores[ref]
        distinct[1:] &= s[1:] != s[:-1]
        distinct[:-1] &= s[:-1] != s[1:]
        # the element after position 49 could tie with it; drop the last to be safe
        distinct[-1] = False
        assert np.array_equal(ours[distinct], ref[distinct])
        refcut = g["cuts"][n]; refcut = refcut[refcut >= 0]
        ourcut = np.array(graph.simple_cuts(cube, adj.copy()))
        assert sorted(ourcut.tolist()) == sorted(refcut.tolist())
        a0 = adj.copy(); np.fill_diagonal(a0, 0)
        contains = np.where(cube == 1)[0]
        cs = np.zeros(c); cs[contains] = a0[contains][:, contains].sum(0)
        assert np.array_equal(cs[ourcut], cs[refcut])


def test_known_answer_3x4():
    # 3 cubes x 4 cards, card 3 never appears
    x = np.array([[1, 1, 0, 0], [1, 0, 1, 0], [1, 1, 0, 0]], dtype=np.float64)
    m = graph.create_adjacency_matrix(x)
    expect = np.array([[1, 2 / 3, 1 / 3, 0], [1, 1, 0, 0], [1, 0, 1, 0], [0, 0, 0, 0]])
    assert np.array_equal(m, expect)
    ip, ix = dense_to_csr(x)
    assert np.array_equal(graph.cooc_counts(ip, ix, 4),
                          np.array([[3, 2, 1, 0], [2, 2, 0, 0], [1, 0, 1, 0], [0, 0, 0, 0]]))
    mh = graph.m_hat(m)
    assert np.allclose(mh.sum(1), 1.0)
    assert np.array_equal(mh[3], np.array([0, 0, 0, 1.0]))   # unseen card -> e_i
    ns = graph.neg_sampler(mh)
    assert abs(ns.sum() - 1) < 1e-15


def test_duplicate_card_ids_collapse():
    ip = np.array([0, 3, 5]); ix = np.array([1, 1, 2, 0, 1], dtype=np.int32)
    cnt = graph.cooc_counts(ip, ix, 3)
    assert np.array_equal(cnt, np.array([[1, 1, 0], [1, 2, 1], [0, 1, 1]]))


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 40), st.integers(1, 60), st.integers(0, 10_000))
def test_property_loop_equals_xtx(k, c, seed):
    rng = np.random.default_rng(seed)
    x = (rng.random((k, c)) < 0.2).astype(np.float64)
    a = graph.create_adjacency_matrix_loop(x)
    b = graph.create_adjacency_matrix(x)
    assert np.array_equal(a, b)
    ip, ix = dense_to_csr(x)
    cnt = graph.cooc_counts(ip, ix, c)
    assert np.array_equal(cnt, cnt.T)
    assert set(np.unique(np.diagonal(a))) <= {0.0, 1.0}


def test_synth_cubes_shape_and_law():
    ip, ix = synth_cubes_csr(64, 3000, cfg=1)
    sizes = np.diff(ip)
    assert sizes.min() >= 360 and sizes.max() <= 720
    for r in range(64):
        row = ix[ip[r]:ip[r + 1]]
        assert np.all(np.diff(row) > 0)            # sorted, distinct
    ip2, ix2 = synth_cubes_csr(64, 3000, cfg=1)
    assert np.array_equal(ix, ix2)                 # seeded
    # heavy tail: low ids far more frequent than high ids
    assert (ix < 300).mean() > 3 * (ix >= 2700).mean()


def test_golden_pairwise_recs(pairwise_golden):
    """Cubes larger than 128 cards: NumPy's recursive pairwise sum is part of what
    the reference ranks (recommend.py:10-13 sums an F-ordered fancy-index copy)."""
    g = pairwise_golden
    c = int(g["num_cards"])
    dense = csr_to_dense(g["indptr"], g["indices"], c)
    adj = graph.create_adjacency_matrix(dense)
    for n, row in enumerate(g["rec_cube_rows"]):
        cube = dense[row]
        scores = graph.simple_recs_scores(cube, adj)
        assert np.array_equal(scores, g["rec_scores"][n])
        ours = np.array(graph.simple_recs(cube, adj))[:50]
        assert np.array_equal(scores[ours], scores[g["recs_top50"][n]])
        refcut = g["cuts"][n]; refcut = refcut[refcut >= 0]
        ourcut = np.array(graph.simple_cuts(cube, adj.copy()))
        cs = g["cut_scores"][n]
        assert np.array_equal(cs[ourcut], cs[refcut])


@pytest.mark.parametrize("k,c,block", [(300, 1000, 384), (64, 130, 50), (500, 257, 4096)])
def test_blocked_dense_counts_equal_sparse_counts(k, c, block):
    """oracle.graph.cooc_counts_blocked (the form used for the BASELINE-sized GPU parity test: K = 20 000, C = 21 000)
    equals the pinned cooc_counts, including duplicate card ids inside a cube."""
    from cubecobrarecommender_b200.synth import synth_cubes_csr
    ip, ix = synth_cubes_csr(k, c, size_lo=1, size_hi=min(c, 120), seed=k + c)
    ix = ix.copy(); ix[1] = ix[0]                              # a duplicate id in the first cube
    a = graph.cooc_counts(ip, ix, c)
    b = graph.cooc_counts_blocked(ip, ix, c, block=block)
    assert b.dtype == np.int32 and np.array_equal(a, b)


def test_m_hat_rows_equals_full_m_hat():
    from cubecobrarecommender_b200.synth import synth_cubes_csr
    k, c = 120, 300
    ip, ix = synth_cubes_csr(k, c, size_lo=3, size_hi=40, seed=5)
    full = graph.m_hat(graph.adjacency_from_counts(graph.cooc_counts(ip, ix, c)))
    assert (np.diagonal(graph.cooc_counts(ip, ix, c)) == 0).any()          # unseen cards exercise the e_i rows
    rows = np.array([0, 7, 7, 299, 150] + list(np.where(np.diagonal(graph.cooc_counts(ip, ix, c)) == 0)[0][:3]))
    assert np.array_equal(graph.m_hat_rows(ip, ix, c, rows), full[rows])
