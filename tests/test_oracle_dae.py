"""oracle/dae.py: the two independent formulations (NumPy float64 hand-derived gradients
vs torch-CPU autograd) agree, plus hand-computed micro cases of the Keras-2.5
conventions (reference src/ml/model.py:20-125, src/ml/train.py:83-88).
PARITY UNPINNED vs TensorFlow (not installable) -- see oracle/dae.py header."""
import numpy as np
import torch

from oracle import dae, graph
from cubecobrarecommender_b200.synth import synth_cubes_csr, csr_to_dense


def _setup(c=96, k=40, b=12, seed=3):
    ip, ix = synth_cubes_csr(k, c, size_lo=6, size_hi=30, seed=seed)
    dense = csr_to_dense(ip, ix, c)
    adj = graph.create_adjacency_matrix(dense)
    mh = graph.m_hat(adj)
    rng = np.random.default_rng(seed)
    x = dense[:b].copy()
    y = x.copy()
    # a fixed "noise" outcome: drop two cards from x (one of them from y), add one
    for r in range(b):
        inc = np.where(x[r] == 1)[0]; exc = np.where(x[r] == 0)[0]
        x[r, inc[:2]] = 0; y[r, inc[0]] = 0; x[r, exc[rng.integers(len(exc))]] = 1
    reg_rows = rng.integers(0, c, size=b)
    params = dae.init_params(c, seed=0)
    # non-zero biases so bias gradients matter
    for kname in params:
        if kname.endswith("bias"):
            params[kname] = (rng.standard_normal(params[kname].shape) * 0.05).astype(np.float32)
    return params, x, y, reg_rows, mh[reg_rows]


def test_param_count_matches_shipped_checkpoint():
    # 1538*C + 518848 parameters, C = 20884 -> 32 638 440
    n = sum(fi * fo + fo for _, fi, fo in dae.layer_specs(20884))
    assert n == 32_638_440


def test_numpy_vs_torch_autograd_float64():
    params, x, y, reg_rows, t = _setup()
    (tot, bce, kl), grads = dae.loss_and_grads_np(params, x, y, reg_rows, t, reg=0.1)
    tm = dae.TorchDAE(params, dtype=torch.float64)
    total, bce_t, kl_t = tm.loss(torch.tensor(x), torch.tensor(y), torch.tensor(reg_rows),
                                 torch.tensor(t), 0.1)
    total.backward()
    assert abs(float(total) - tot) < 1e-12 and abs(float(bce_t) - bce) < 1e-12
    assert abs(float(kl_t) - kl) < 1e-12
    tg = tm.grads()
    for kname, g in grads.items():
        scale = np.abs(g).max() + 1e-30
        assert np.abs(tg[kname] - g).max() / scale < 1e-9, kname


def test_kl_clip_gradient_mask():
    """q below 1e-7 is clipped: no gradient flows through those entries (S excludes them)."""
    z = np.array([[0.0, -30.0, 1.0]])
    t = np.array([[0.5, 0.0, 0.5]])
    q = dae.softmax_np(z)
    assert q[0, 1] < 1e-7
    zt = torch.tensor(z, requires_grad=True)
    qt = torch.softmax(zt, 1)
    tc = torch.clamp(torch.tensor(t), 1e-7, 1.0)
    loss = (tc * torch.log(tc / torch.clamp(qt, 1e-7, 1.0))).sum(1).mean()
    loss.backward()
    tcn = np.clip(t, 1e-7, 1)
    unclipped = q >= 1e-7
    s = (tcn * unclipped).sum(1, keepdims=True)
    expect = q * s - tcn * unclipped
    assert np.allclose(zt.grad.numpy(), expect, atol=1e-15)
    assert abs(dae.kld_np(t, q) - float(loss)) < 1e-15


def test_bce_logits_micro_case():
    z = np.array([[0.0, 2.0], [-3.0, 50.0]]); y = np.array([[1.0, 0.0], [0.0, 1.0]])
    hand = np.mean([np.log(2), 2 + np.log1p(np.exp(-2)), np.log1p(np.exp(-3)), np.log1p(np.exp(-50.0))])
    assert abs(dae.bce_from_logits_np(z, y) - hand) < 1e-15


def test_adam_tf_style_first_steps():
    p = {"w": np.array([1.0, -2.0])}; g = {"w": np.array([0.5, -0.25])}
    m = {"w": np.zeros(2)}; v = {"w": np.zeros(2)}
    dae.adam_step_np(p, g, m, v, 1)
    # step 1: m = .1 g, v = .001 g^2, lr_t = 1e-3*sqrt(.001)/.1
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expect = np.array([1.0, -2.0]) - lr_t * (0.1 * g["w"]) / (np.sqrt(0.001 * g["w"] ** 2) + 1e-7)
    assert np.allclose(p["w"], expect, rtol=0, atol=1e-15)


def test_float32_train_steps_track_float64():
    params, x, y, reg_rows, t = _setup()
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p64.items()}; v_ = {k: np.zeros_like(v) for k, v in p64.items()}
    tm = dae.TorchDAE(params)
    xt, yt = torch.tensor(x, dtype=torch.float32), torch.tensor(y, dtype=torch.float32)
    rt, tt = torch.tensor(reg_rows), torch.tensor(t, dtype=torch.float32)
    for step in range(1, 4):
        (tot, _, _), grads = dae.loss_and_grads_np(p64, x, y, reg_rows, t, reg=0.1)
        dae.adam_step_np(p64, grads, m, v_, step)
        tot32, _, _ = tm.train_step(xt, yt, rt, tt, 0.1)
        assert abs(tot32 - tot) / abs(tot) < 1e-5


def test_rank_additions_tie_rule():
    res = np.array([0.5, 0.9, 0.9, 0.1, 0.9]); in_cube = np.array([0, 0, 1, 0, 0])
    assert dae.rank_additions(res, in_cube, 3) == [4, 1, 0]


def test_similarity_restatement_known_answers():
    """Keras CosineSimilarity loss as the reference's similarity.py uses it: -cos, l2_normalize with eps 1e-12."""
    from oracle import dae as od
    embs = np.array([[3.0, 4.0], [6.0, 8.0], [4.0, -3.0], [0.0, 0.0], [-3.0, -4.0]])
    d = od.similarity_np(embs, 0)
    assert np.allclose(d, [-1.0, -1.0, 0.0, 0.0, 1.0])           # itself, parallel, orthogonal, zero vector, opposite
    assert d.argsort(kind="stable")[0] == 0
    # card_embeddings_np == encoder applied to the dense identity (what the reference feeds Keras)
    params = od.init_params(40, seed=1)
    eye = np.eye(40)
    p = {k: v.astype(np.float64) for k, v in params.items()}
    h = eye
    for n in ("encoder_e1", "encoder_e2", "encoder_e3", "encoder_bottleneck"):
        h = np.maximum(h @ p[n + "/kernel"] + p[n + "/bias"], 0.0)
    assert np.allclose(od.card_embeddings_np(params), h, atol=1e-14)


def test_onehot_rows_as_gather_is_the_same_arithmetic():
    """The BASELINE-sized parity cases evaluate the second tower's first layer as a row gather of its kernel (and its
    gradient as a row scatter-add) instead of multiplying by one-hot rows of I: identical float64 numbers."""
    params, x, y, reg_rows, t = _setup()
    reg_rows = reg_rows.copy(); reg_rows[1] = reg_rows[0]            # a row drawn twice
    t = t.copy(); t[1] = t[0]
    (la, ga) = dae.loss_and_grads_np(params, x, y, reg_rows, t, reg=0.1)
    (lb, gb) = dae.loss_and_grads_np(params, x, y, reg_rows, t, reg=0.1, onehot_rows_as_gather=True)
    assert la == lb
    for kname in ga:
        assert np.abs(ga[kname] - gb[kname]).max() <= 1e-18 + 1e-15 * np.abs(ga[kname]).max(), kname


def test_published_keras_known_answers():
    """The worked examples of the Keras API documentation (tf.keras.losses.BinaryCrossentropy / KLDivergence,
    tf.keras.metrics.BinaryAccuracy / CategoricalAccuracy, tf.keras.optimizers.Adam -- the figures printed there, three
    digits) against the oracle's restatement.  TensorFlow 2.5.2 (requirements.txt:48) is not installable here, so these
    published numbers are the only vectors of the dependency itself the oracle can be held to."""
    y_true = np.array([[0.0, 1.0], [0.0, 0.0]]); y_pred = np.array([[0.6, 0.4], [0.4, 0.6]])
    # BinaryCrossentropy()(y_true, y_pred) = 0.815 (docs); the oracle works from logits
    z = np.log(y_pred / (1 - y_pred))
    assert abs(dae.bce_from_logits_np(z, y_true) - 0.815) < 1e-3
    per_row = [dae.bce_from_logits_np(z[i:i + 1], y_true[i:i + 1]) for i in range(2)]
    assert abs(per_row[0] - 0.916) < 1e-3 and abs(per_row[1] - 0.714) < 1e-3          # reduction=NONE in the docs
    # KLDivergence()(y_true, y_pred) = 0.458, per row [0.916, -3.08e-06] (docs): y_true is clipped to [1e-7, 1]
    assert abs(dae.kld_np(y_true, y_pred) - 0.458) < 1e-3
    assert abs(dae.kld_np(y_true[1:2], y_pred[1:2]) - (-3.08e-06)) < 1e-8
    # BinaryAccuracy: y_true [[1],[1],[0],[0]], y_pred [[0.98],[1],[0],[0.6]] -> 0.75 (docs); logits of the predictions
    yt = np.array([[1.0], [1.0], [0.0], [0.0]]); yp = np.array([[0.98], [1.0 - 1e-12], [1e-12], [0.6]])
    assert dae.binary_accuracy_np(np.log(yp / (1 - yp)), yt) == 0.75
    # CategoricalAccuracy: y_true [[0,0,1],[0,1,0]], y_pred [[0.1,0.9,0.8],[0.05,0.95,0]] -> 0.5 (docs)
    ct = np.array([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0]]); cp = np.array([[0.1, 0.9, 0.8], [0.05, 0.95, 0.0]])
    assert dae.categorical_accuracy_np(cp, ct) == 0.5                                   # (argmax of logits = argmax of probabilities)
    # Adam(learning_rate=0.1): var = 10.0, loss = var^2 / 2 -> after one step var = 9.9 (docs)
    p = {"w": np.array([10.0])}; m = {"w": np.zeros(1)}; v = {"w": np.zeros(1)}
    dae.adam_step_np(p, {"w": p["w"].copy()}, m, v, 1, lr=0.1)
    assert abs(p["w"][0] - 9.9) < 1e-6
