"""Drop-in surface on the GPU: create_mtx files, simple_recs/simple_cuts signatures, ML recommendation
(ids bit-exact vs the oracle ranking of the same scores), DataGenerator mirror, fit(), save/load."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from cubecobrarecommender_b200.ml import generator as GEN, inference as INF, model as M, train as T
from cubecobrarecommender_b200.non_ml import create_mtx, utils
from cubecobrarecommender_b200.scripts import cut_cards, recommend
from cubecobrarecommender_b200.sparse import CubeCSR
from cubecobrarecommender_b200.synth import csr_to_dense, synth_cubes_csr
from oracle import dae as od, graph as og, noise as on


def _write_dataset(root, k=60, c=90, seed=5):
    ip, ix = synth_cubes_csr(k, c, size_lo=5, size_hi=30, seed=seed)
    names = {f"card {i}": [f"id{i}a", f"id{i}b"] for i in range(c)}
    os.makedirs(root / "data/maps"); os.makedirs(root / "data/cube")
    json.dump(names, open(root / "data/maps/nameToId.json", "w"))
    cubes = [{"_id": f"c{r}", "cards": [{"cardID": f"id{j}{'ab'[j % 2]}"} for j in ix[ip[r]:ip[r + 1]]]} for r in range(k)]
    json.dump(cubes[:30], open(root / "data/cube/a.json", "w"))
    json.dump(cubes[30:], open(root / "data/cube/b.json", "w"))
    return c


def test_create_mtx_cli_files(tmp_path, capsys):
    c = _write_dataset(tmp_path)
    adj = create_mtx.main(str(tmp_path))
    out = capsys.readouterr().out
    assert out.startswith("getting data\ncreating matrix\n1 / 90\n")
    saved = np.load(tmp_path / "output/full_adj_mtx.npy")
    assert saved.dtype == np.float64 and saved.shape == (c, c) and np.array_equal(saved, adj)
    i2c = json.load(open(tmp_path / "output/int_to_card.json"))
    assert i2c["0"] == "card 0" and len(i2c) == c
    # same answer as the oracle on the dense cubes the reference loader builds
    n, lookup, c2i, _ = utils.get_card_maps(str(tmp_path / "data/maps/nameToId.json"))
    dense = utils.build_cubes(str(tmp_path / "data/cube"), utils.get_num_cubes(str(tmp_path / "data/cube")), n, lookup, c2i)
    assert np.array_equal(saved, og.create_adjacency_matrix(dense))
    assert np.array_equal(utils.create_adjacency_matrix(dense, verbose=False, force_diag=0.0),
                          og.create_adjacency_matrix(dense, force_diag=0.0))


def test_simple_recs_and_cuts_dropin(pairwise_golden):
    g = pairwise_golden
    c = int(g["num_cards"])
    dense = csr_to_dense(g["indptr"], g["indices"], c)
    adj = og.create_adjacency_matrix(dense)
    i2c = {i: f"n{i}" for i in range(c)}
    for n, row in enumerate(g["rec_cube_rows"]):
        cube = dense[row]
        full = recommend.simple_recs(cube, adj)
        assert full == [int(i) for i in og.simple_recs(cube, adj)]              # whole ranking, ids exact
        assert len(full) == int((cube == 0).sum())
        assert recommend.simple_recs(cube, adj, i2c)[:5] == [i2c[i] for i in full[:5]]
        a2 = adj.copy()
        cuts = cut_cards.simple_cuts(cube, a2)
        assert (np.diagonal(a2) == 0).all()                                      # mutates like the reference
        assert cuts == [int(i) for i in og.simple_cuts(cube, adj.copy())]
        s = g["rec_scores"][n]
        assert np.array_equal(s[full[:50]], s[g["recs_top50"][n]])              # the reference's own answer


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_ml_recommend_batched(precision):
    c, k = 400, 37
    ip, ix = synth_cubes_csr(k, c, size_lo=0, size_hi=120, seed=2)
    lists = [ix[ip[i]:ip[i + 1]] for i in range(k)]; lists[3] = np.zeros(0, np.int32)
    csr = CubeCSR.from_lists(lists, c)
    params = od.init_params(c, seed=4)
    rng = np.random.default_rng(0)
    for kk in params:      # larger weights -> a wide spread of probabilities, some saturated at exactly 1.0
        params[kk] = (params[kk] * 2 + (rng.standard_normal(params[kk].shape) * 0.1 if kk.endswith("bias") else 0)).astype(np.float32)
    model = M.CC_Recommender(c, device="cuda", precision=precision)
    model.set_weights_dict(params)
    rec = INF.MLRecommender(model, chunk=16)
    probs = rec.probabilities(csr).cpu().numpy()
    ref = od.recommend_np(params, csr.to_dense())
    assert np.abs(probs - ref).max() < (1e-4 if precision == "fp32" else 2e-2)
    ids, vals, cnt = rec.recommend(csr, 50)
    dense = csr.to_dense()
    for r in range(k):
        expect = od.rank_additions(probs[r], dense[r], 50)          # oracle ranking of the SAME scores
        assert cnt[r] == len(expect)
        assert ids[r, :cnt[r]].tolist() == expect                   # ids bit-exact, ties included
        assert np.array_equal(vals[r, :cnt[r]], probs[r][expect])
    ids3, vals3, cnt3 = rec.recommend(csr.pin_memory(), 50, copy=False)      # pinned batch in, pinned views out
    assert np.array_equal(ids3, ids) and np.array_equal(vals3, vals) and np.array_equal(cnt3, cnt)
    ids2, vals2, cnt2 = rec.recommend(csr, 200)      # n > 128: sigmoid pass + radix select instead of the fused select
    for r in range(k):
        expect = od.rank_additions(probs[r], dense[r], 200)
        assert cnt2[r] == len(expect) and ids2[r, :cnt2[r]].tolist() == expect
    i2c = {i: f"card{i}" for i in range(c)}
    out = rec.recommend_one([int(v) for v in lists[5]], 7, i2c)
    assert list(out) == ["additions", "cuts"] and len(out["additions"]) == 7
    assert set(out["cuts"]) == {i2c[int(v)] for v in lists[5]}
    assert list(out["additions"]) == [i2c[i] for i in od.rank_additions(probs[5], dense[5], 7)]


def test_web_get_ml_recommend_resident(tmp_path, monkeypatch):
    from cubecobrarecommender_b200.web import ml_recommend_web as W
    c = 120
    model = M.CC_Recommender(c, device="cuda", seed=3)
    model.save(str(tmp_path / "ml_files/recommender"))
    json.dump({str(i): f"card {i}" for i in range(c)}, open(tmp_path / "idmap.json", "w"))
    monkeypatch.setenv("CUBECOBRA_MODEL_DIR", str(tmp_path / "ml_files/recommender"))
    monkeypatch.setenv("CUBECOBRA_ID_MAP", str(tmp_path / "idmap.json"))
    monkeypatch.setitem(W._state, "rec", None)
    monkeypatch.setitem(W._state, "batcher", None)
    names = ["Card 5", "CARD 17", "not a card", "card 99"]
    out = W.get_ml_recommend("ignored", 10, card_names=names)
    assert len(out["additions"]) == 10 and set(out["cuts"]) == {"card 5", "card 17", "card 99"}
    assert not set(out["additions"]) & set(out["cuts"])
    assert all(isinstance(v, float) and 0 <= v <= 1 for v in out["additions"].values())
    loaded = M.load_model(str(tmp_path / "ml_files/recommender"))
    a, b = model.get_weights_dict(), loaded.get_weights_dict()
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_web_requests_are_micro_batched(monkeypatch):
    """concurrent get_ml_recommend calls (request threads of the WSGI server) are collected by the
    RequestBatcher and served by ONE batched recommend_device call; every request still gets exactly the answer the
    one-cube path gives (same ids in rank order, same scores, cuts in cubelist order), whatever it was batched with."""
    import threading
    from cubecobrarecommender_b200.web import ml_recommend_web as W
    c, k = 400, 24
    ip, ix = synth_cubes_csr(k, c, size_lo=1, size_hi=120, seed=12)
    model = M.CC_Recommender(c, device="cuda", seed=5, precision="tf32")
    rec = INF.MLRecommender(model, chunk=64)
    i2c = {i: f"card {i}" for i in range(c)}
    saved = dict(W._state)
    batcher = W.install(rec, i2c, max_batch=64, max_wait_ms=200.0)
    try:
        cubes = []
        for r in range(k):
            idx = [int(v) for v in ix[ip[r]:ip[r + 1]]]
            rng = np.random.default_rng(r)
            rng.shuffle(idx)                                  # cubelist order is not sorted
            if r % 5 == 0 and idx:
                idx.append(idx[0])                            # a card listed twice
            cubes.append(idx)
        cubes[3] = []                                         # an empty cube list (all cards unknown)
        amounts = [7 + (r % 4) * 50 for r in range(k)]        # 7, 57, 107, 157: fused and unfused select in one batch
        outs = [None] * k
        gate = threading.Barrier(k)

        def request(r):
            gate.wait()
            outs[r] = W.get_ml_recommend("ignored", amounts[r], card_names=[f"Card {i}" for i in cubes[r]])
        threads = [threading.Thread(target=request, args=(r,)) for r in range(k)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
        assert all(o is not None for o in outs)
        assert batcher.stats["requests"] == k and batcher.stats["batches"] < k and batcher.stats["max_batch"] > 1
        for r in range(k):
            one = rec.recommend_one(cubes[r], amounts[r], i2c)
            assert list(outs[r]["additions"]) == list(one["additions"]), r           # same cards, same order
            assert list(outs[r]["cuts"]) == list(one["cuts"]), r
            a = np.array(list(outs[r]["additions"].values())); b = np.array(list(one["additions"].values()))
            assert np.abs(a - b).max() < 1e-6 if len(a) else True    # (a different GEMM batch shape: last-bit differences)
            assert len(outs[r]["additions"]) == min(amounts[r], c - len(set(cubes[r])))
        # a lone request is served too (the worker does not wait for company beyond max_wait_ms)
        solo = W.get_ml_recommend("ignored", 5, card_names=["card 1", "card 2"])
        assert len(solo["additions"]) == 5 and set(solo["cuts"]) == {"card 1", "card 2"}
        assert W.get_ml_recommend("ignored", 3, non_json=True, card_names=["card 1"]) is None
    finally:
        batcher.close()
        W._state.update(saved)
        W._state["batcher"] = None          # (install() closed whatever batcher was there before)


def test_datagenerator_mirror_and_fit():
    c, k, b = 256, 96, 32
    ip, ix = synth_cubes_csr(k, c, size_lo=10, size_hi=60, seed=8)
    dense = csr_to_dense(ip, ix, c)
    mh = og.m_hat(og.create_adjacency_matrix(dense))
    np.random.seed(0)
    gen = GEN.DataGenerator(mh, dense, batch_size=b, noise=0.2, seed=11)
    assert len(gen) == 3 and gen.N_cubes == k and gen.N_cards == c
    assert np.abs(gen.neg_sampler - og.neg_sampler(mh)).max() < 1e-15
    (x, xr), (y, yr) = gen[1]
    assert x.shape == (b, c) and x.dtype == np.float64 and xr.shape == (b, c) and yr.shape == (b, c)
    assert (xr.sum(1) == 1).all() and np.array_equal(yr, mh[np.argmax(xr, 1)])
    on.check_noise_invariants(dense[gen.indices[b:2 * b]], x, y)
    model = M.CC_Recommender(c, device="cuda", seed=0, precision="tf32")
    hist = T.fit(model, gen, epochs=6, reg=0.1, log=lambda *_: None)
    assert len(hist) == 6 and hist[-1]["loss"] < hist[0]["loss"]                 # it trains
    assert 0.0 <= hist[0]["output_2_accuracy"] <= 1.0 and hist[-1]["output_1_accuracy"] >= hist[0]["output_1_accuracy"] - 0.05
    assert set(hist[0]) == {"loss", "output_1_loss", "output_2_loss", "output_1_accuracy", "output_2_accuracy"}
    assert abs(hist[0]["loss"] - (hist[0]["output_1_loss"] + 0.1 * hist[0]["output_2_loss"])) < 1e-9
    assert int(model.store.step.item()) == 18


def test_fit_periodic_save_and_resume(tmp_path):
    """a run interrupted after a periodic checkpoint and resumed from it ends where the uninterrupted
    run ends -- the epoch shuffle is a function of (seed, epoch), the noise draws of (seed, Adam step), both restored
    from the checkpoint together with the weights and the Adam slots."""
    c, k, b = 256, 96, 32
    ip, ix = synth_cubes_csr(k, c, size_lo=10, size_hi=60, seed=8)
    dense = csr_to_dense(ip, ix, c)
    mh = og.m_hat(og.create_adjacency_matrix(dense))
    quiet = lambda *_: None

    def run(epochs, model, initial_epoch=0, ckpt=None, save_every=None):
        gen = GEN.DataGenerator(mh, dense, batch_size=b, noise=0.2, seed=11)
        return T.fit(model, gen, epochs=epochs, reg=0.1, log=quiet, initial_epoch=initial_epoch, checkpoint_dir=ckpt,
                     save_every=save_every), gen
    straight = M.CC_Recommender(c, device="cuda", seed=0, precision="fp32")
    h_all, gen_all = run(4, straight)
    ck = str(tmp_path / "ml_files/run")
    first = M.CC_Recommender(c, device="cuda", seed=0, precision="fp32")
    h_first, _ = run(3, first, ckpt=ck, save_every=2)             # checkpoint after epoch 2, "crash" after epoch 3
    resumed = M.CC_Recommender.load(ck, device="cuda", precision="fp32")
    assert resumed.completed_epochs == 2 and int(resumed.store.step.item()) == 2 * (k // b)
    h_rest, gen_rest = run(4, resumed, initial_epoch=resumed.completed_epochs)
    assert len(h_rest) == 2
    assert np.array_equal(gen_rest.indices, gen_all.indices)       # same permutation at the same epoch
    # exact-fp32 mode: only float atomics (first-layer scatter-add) separate the two runs
    for a, bb in zip(h_all[2:], h_rest):
        assert abs(a["loss"] - bb["loss"]) < 1e-5 * abs(a["loss"])
    wa, wb = straight.get_weights_dict(), resumed.get_weights_dict()
    assert max(np.abs(wa[n] - wb[n]).max() for n in wa) < 2e-4
    assert int(resumed.store.step.item()) == int(straight.store.step.item()) == 4 * (k // b)
    # two generators with the same seed agree epoch by epoch (what keeps data-parallel ranks on the same permutation)
    g1 = GEN.DataGenerator(mh, dense, batch_size=b, seed=3); g2 = GEN.DataGenerator(mh, dense, batch_size=b, seed=3)
    for _ in range(3):
        assert np.array_equal(g1.indices, g2.indices) and sorted(g1.indices.tolist()) == list(range(k))
        prev = g1.indices.copy(); g1.on_epoch_end(); g2.on_epoch_end()
        assert not np.array_equal(prev, g1.indices)


def test_noise_kernel_overflow_leaves_an_empty_row():
    """A cube larger than the engine's max_cube_size sets the overflow flag AND leaves x empty / y zero for that row
    (never the previous batch's data); fit() and the host-fed stream check the flag."""
    from cubecobrarecommender_b200 import graph as G
    from cubecobrarecommender_b200.ml import engine as E
    c, b = 384, 8
    lists = [list(range(0, 20 + 3 * i)) for i in range(b)]
    lists[5] = list(range(0, 200))                                 # larger than max_cube_size = 64
    csr = CubeCSR.from_lists(lists, c)
    gr = G.build_graph(csr, "cuda", want_m64=False)
    prob, alias = E.alias_table(gr.neg_sampler.cpu().numpy(), "cuda")
    model = M.CC_Recommender(c, device="cuda", precision="tf32")
    eng = E.DAEEngine(model, gr.mhat, batch=b, reg_rows=b, reg=0.1, max_cube_size=64)
    eng.x_len.fill_(7); eng.y_bits.fill_(-1); eng.x_dense.fill_(1.0)
    indptr, indices = G.upload_csr(csr, "cuda")
    eng.sample_batch(indptr, indices, None, prob, alias, seed=1)
    assert int(eng.overflow.item()) == 1
    assert int(eng.x_len[5].item()) == 0 and (eng.y_bits[5] == 0).all() and (eng.x_dense[5] == 0).all()
    assert int(eng.x_len[4].item()) > 0 and (eng.x_dense[4].sum() == eng.x_len[4]).item()
    with pytest.raises(RuntimeError, match="overflow"):
        eng.check_overflow()


def test_similarity_script_and_kernel_vs_oracle(tmp_path, capsys):
    """scripts/similarity.py (reference src/scripts/similarity.py): encoder embeddings of all one-hot cards, Keras
    CosineSimilarity loss against one card, ascending argsort.  Checked against the float64 restatement; ids must
    match wherever the oracle's distances are separated by more than the float32 tolerance."""
    from cubecobrarecommender_b200.scripts import similarity as SIM
    c = 300
    params = od.init_params(c, seed=3)
    model = M.CC_Recommender(c, device="cuda", precision="fp32")
    model.set_weights_dict(params)
    model.save(str(tmp_path / "ml_files/high_req"))
    i2c = {str(i): f"card {i}" for i in range(c)}
    json.dump(i2c, open(tmp_path / "id_map.json", "w"))
    rec = INF.MLRecommender(model)
    emb = rec.card_embeddings().cpu().numpy()
    ref_emb = od.card_embeddings_np(params)
    assert np.abs(emb - ref_emb).max() < 1e-5
    for q in (0, 17, 299):
        ids, dists = rec.similar(q, 25)
        ref = od.similarity_np(ref_emb, q)
        assert ids[0] == q and abs(dists[0] + 1.0) < 1e-6                  # the card itself: cosine 1
        assert np.abs(dists - ref[ids]).max() < 1e-5
        order = ref.argsort(kind="stable")[:25]
        gaps_ok = np.abs(np.diff(ref[order])) > 1e-5                       # well separated neighbours
        same = ids == order
        assert same[1:-1][gaps_ok[:-1] & gaps_ok[1:]].all()                # separated from both neighbours
        assert (np.diff(dists) >= 0).all()
    out = SIM.main(["card_17", "5"], model_dir=str(tmp_path / "ml_files/high_req"), id_map=str(tmp_path / "id_map.json"))
    printed = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 5 and out[0][0] == "card 17" and printed[0].startswith("1: card 17 -1")
