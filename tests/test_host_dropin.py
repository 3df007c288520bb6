"""Host-side drop-in logic that needs no GPU: loaders (reference src/non_ml/utils.py:6-73), name
normalisation (recommend.py:53), the WSGI route's error strings (web/__init__.py:18-30), CSR helpers."""
import json
import os

import numpy as np
import pytest

from cubecobrarecommender_b200.non_ml import utils
from cubecobrarecommender_b200.scripts import common
from cubecobrarecommender_b200.sparse import CubeCSR


@pytest.fixture
def data_dir(tmp_path):
    names = {"alpha": ["a1", "a2"], "beta": ["b1"], "gamma": ["g1"], "plains": ["p1"], "delta": ["d1"]}
    os.makedirs(tmp_path / "data/maps"); os.makedirs(tmp_path / "data/cube")
    json.dump(names, open(tmp_path / "data/maps/nameToId.json", "w"))
    cubes_a = [{"_id": "c0", "cards": [{"cardID": "a1"}, {"cardID": "b1"}, {"cardID": "a2"}, {"cardID": "zzz"}]},
               {"_id": "c1", "cards": [{"cardID": "g1"}]}]
    cubes_b = [{"_id": "c2", "cards": []}, {"_id": "c3", "cards": [{"cardID": "d1"}, {"cardID": "p1"}, {"cardID": "b1"}]}]
    json.dump(cubes_a, open(tmp_path / "data/cube/0.json", "w"))
    json.dump(cubes_b, open(tmp_path / "data/cube/1.json", "w"))
    return tmp_path


def test_loaders_match_reference_semantics(data_dir):
    n, name_lookup, card_to_int, int_to_card = utils.get_card_maps(str(data_dir / "data/maps/nameToId.json"))
    assert n == 5 and card_to_int["alpha"] == 0 and card_to_int["delta"] == 4       # JSON insertion order
    assert name_lookup["a2"] == "alpha" and int_to_card[2] == "gamma"
    k = utils.get_num_cubes(str(data_dir / "data/cube"))
    assert k == 4
    dense = utils.build_cubes(str(data_dir / "data/cube"), k, n, name_lookup, card_to_int)
    csr = utils.build_cubes_csr(str(data_dir / "data/cube"), n, name_lookup, card_to_int)
    assert dense.dtype == np.float64 and dense.shape == (4, 5)
    assert np.array_equal(csr.to_dense(), dense)                 # same cubes, same (listdir) order
    assert dense.sum() == 6                                       # duplicate ids collapse, unknown ids skipped
    # with an exclude file the basic lands drop out (utils.py:35-36)
    json.dump({"x": {"isToken": True, "name_lower": "gamma"}}, open(data_dir / "cards.json", "w"))
    n2, _, c2i, _ = utils.get_card_maps(str(data_dir / "data/maps/nameToId.json"), str(data_dir / "cards.json"))
    assert n2 == 3 and "plains" not in c2i and "gamma" not in c2i


def test_normalise_and_cube_indices():
    assert common.normalise("Lim-Dûl's Vault") == "lim-dul's vault"
    assert common.normalise("Æther Vial") == "aether vial"
    c2i = {"lim-dul's vault": 3, "island": 0}
    assert common.cube_indices(["Lim-Dûl's Vault", "Custom Card", "Island", ""], c2i) == [3, 0]
    v = common.cube_vector([3, 0, 3], 5)
    assert v.tolist() == [1, 0, 0, 1, 0]


def test_int_to_card_json_roundtrip(tmp_path):
    json.dump({0: "a", 1: "b"}, open(tmp_path / "m.json", "w"))          # int keys become strings (create_mtx.py:28-29)
    i2c, c2i = common.load_int_to_card(str(tmp_path / "m.json"))
    assert i2c == {0: "a", 1: "b"} and c2i == {"a": 0, "b": 1}


def test_wsgi_error_strings():
    from cubecobrarecommender_b200.web import app

    def call(qs):
        out = {}
        body = app({"QUERY_STRING": qs}, lambda s, h: out.update(status=s))
        return out["status"], b"".join(body).decode()
    assert call("")[1] == "Need cube_name and num_recs as parameters!"
    assert call("cube_name=x&num_recs=abc")[1] == "num_recs needs to be an integer!"


def test_csr_rows_and_lists():
    csr = CubeCSR.from_lists([[4, 1, 1], [], [0, 2]], 5)
    assert csr.indices.tolist() == [1, 4, 0, 2] and csr.max_size == 2
    sub = csr.rows([2, 0])
    assert sub.indptr.tolist() == [0, 2, 4] and sub.indices.tolist() == [0, 2, 1, 4]
    with pytest.raises(ValueError):
        CubeCSR.from_lists([[7]], 5)


def test_savedmodel_key_mapping_and_clear_error(tmp_path):
    """The reference's SavedModel directories cannot be read without TensorFlow: load() says so and names the offline
    converter, whose checkpoint-key mapping (pure string logic) covers all 24 tensors, Adam slots and the step."""
    from cubecobrarecommender_b200.scripts import convert_savedmodel as CV
    sfx = CV.SUFFIX
    keys = []
    for (owner, attr) in CV.ATTR_TO_LAYER:
        for p in ("kernel", "bias"):
            keys.append(f"{owner}/{attr}/{p}{sfx}")
            keys.append(f"{owner}/{attr}/{p}/.OPTIMIZER_SLOT/optimizer/m{sfx}")
            keys.append(f"{owner}/{attr}/{p}/.OPTIMIZER_SLOT/optimizer/v{sfx}")
    keys += [f"optimizer/iter{sfx}", f"optimizer/beta_1{sfx}", "_CHECKPOINTABLE_OBJECT_GRAPH", f"keras_api/metrics/0/total{sfx}"]
    m = CV.map_checkpoint_keys(keys)
    names = set(m.values())
    assert len(m) == 24 * 3 + 1 and "step" in names
    assert {"encoder_e1/kernel", "main_reconstruction/bias", "reg_d2/kernel", "adam_m/encoder_bottleneck/bias",
            "adam_v/reg_reconstruction/kernel"} <= names
    assert m[f"decoder_for_reg/reconstruct/kernel{sfx}"] == "reg_reconstruction/kernel"
    # a SavedModel directory: load() fails loudly and points at the converter (torch is imported by the model module)
    pytest.importorskip("torch")
    from cubecobrarecommender_b200.ml.model import CC_Recommender
    d = tmp_path / "recommender"
    (d / "variables").mkdir(parents=True)
    (d / "saved_model.pb").write_bytes(b"\x00")
    with pytest.raises(FileNotFoundError, match="convert_savedmodel"):
        CC_Recommender.load(str(d), device="cpu")


def test_request_batcher_host_logic():
    """web.ml_recommend_web.RequestBatcher with a stand-in recommender (no GPU): concurrent requests are served by one
    batched call, every request gets its own slice (per-request amount, cuts in cubelist order, unknown cards skipped),
    a recommender failure reaches every waiting request, and a dead worker never leaves a request hanging."""
    import threading
    torch = pytest.importorskip("torch")
    from cubecobrarecommender_b200.web import ml_recommend_web as W

    class FakeModel:
        N = 50
        device = torch.device("cpu")

    class FakeRec:
        model = FakeModel()
        fail = False

        def recommend(self, csr, n, want_cuts=False):
            if self.fail:
                raise ValueError("boom")
            k = csr.num_cubes
            ids = np.tile(np.arange(n, dtype=np.int32)[::-1], (k, 1))
            vals = np.tile(np.linspace(1, 0, n, dtype=np.float32), (k, 1))
            cnt = np.array([n - (r % 2) for r in range(k)], dtype=np.int32)
            return ids, vals, cnt, (csr.indices.astype(np.float32) / 100)

    rec = FakeRec()
    saved = dict(W._state)
    b = W.install(rec, {i: f"c{i}" for i in range(50)}, max_batch=8, max_wait_ms=300)
    try:
        outs = [None] * 6
        errs = [None] * 6

        def req(r):
            try:
                outs[r] = W.get_ml_recommend("x", 3 + r, card_names=[f"C{r + 1}", "custom card", f"c{r}", f"c{r}"])
            except Exception as e:
                errs[r] = e
        ts = [threading.Thread(target=req, args=(r,)) for r in range(6)]
        [t.start() for t in ts]
        [t.join(30) for t in ts]
        assert errs == [None] * 6 and b.stats["requests"] == 6 and b.stats["batches"] <= 2 and b.stats["max_batch"] >= 3
        for r in range(6):
            assert list(outs[r]["cuts"]) == [f"c{r + 1}", f"c{r}"]                      # cubelist order, repeats collapse
            assert outs[r]["cuts"][f"c{r}"] == pytest.approx(r / 100)
            assert len(outs[r]["additions"]) >= 3 + r - 1
        rec.fail = True
        with pytest.raises(ValueError, match="boom"):
            W.get_ml_recommend("x", 3, card_names=["c1"])
        rec.fail = False
        assert W.get_ml_recommend("x", 0, card_names=["c1"])["additions"]              # amount <= 0 still yields one card
        b.close()
        with pytest.raises(RuntimeError, match="not running"):
            W.get_ml_recommend("x", 3, card_names=["c1"])
    finally:
        b.close()
        W._state.update(saved)
        W._state["batcher"] = None
