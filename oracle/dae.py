"""Oracle (TEST INFRASTRUCTURE): the regularised denoising auto-encoder step.

CPU restatement of
  * ``Encoder`` / ``Decoder`` / ``CC_Recommender.call``
        -- reference ``src/ml/model.py:20-48, 50-70, 89-125``
  * ``compile(optimizer='adam', loss=['binary_crossentropy',
        'kullback_leibler_divergence'], loss_weights=[1.0, reg])``
        -- reference ``src/ml/train.py:83-88``
  * ``recommend`` (inference) and the ranking walk
        -- reference ``src/scripts/ml_recommend.py:78-116``,
           ``web/ml_recommend_web.py:39-64``

Third-party arithmetic restated here (not present under /root/reference):
TensorFlow / Keras 2.5.2 (reference ``requirements.txt:48``):
  * ``Dense``: ``act(x @ kernel + bias)``, kernel stored ``(in, out)``,
    ``glorot_uniform`` kernels and zero biases;
  * ``binary_crossentropy`` on a sigmoid output inside ``tf.function`` is
    evaluated from the logits: ``max(z,0) - z*y + log1p(exp(-|z|))``, mean over
    the last axis, then mean over the batch (== mean over B*C);
  * ``kullback_leibler_divergence``: ``t' = clip(t, 1e-7, 1)``,
    ``q' = clip(q, 1e-7, 1)``, ``sum_c t' * log(t'/q')``, mean over the batch;
  * ``Adam`` defaults lr=1e-3, b1=0.9, b2=0.999, eps=1e-7 with
    ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t)`` and ``theta -= lr_t*m/(sqrt(v)+eps)``
    (epsilon outside the bias correction), dense updates of every variable.

PARITY UNPINNED for this file: TensorFlow cannot be installed in the build
container and the reference holds no golden outputs, so this restatement is
cross-checked between its two independent formulations (NumPy float64 with
hand-derived gradients vs torch-CPU autograd) in ``tests/test_oracle_dae.py``,
and held to the only vectors of the dependency itself that exist outside it:
the worked examples printed in the Keras API documentation (BinaryCrossentropy
0.815 / [0.916, 0.714], KLDivergence 0.458 / [0.916, -3.08e-06] -- the second
row is the clip of y_true to 1e-7 at work --, BinaryAccuracy 0.75,
CategoricalAccuracy 0.5, one Adam step 10.0 -> 9.9;
``test_published_keras_known_answers``).
"""
from __future__ import annotations

import numpy as np

KERAS_EPS = 1e-7
HIDDEN = (512, 256, 128, 64)


def layer_specs(num_cards: int):
    """(name, fan_in, fan_out) for the 12 Dense layers, in Keras creation
    order (reference ``model.py:27-33, 58-64, 92-98``)."""
    c = num_cards
    enc = [("encoder_e1", c, 512), ("encoder_e2", 512, 256),
           ("encoder_e3", 256, 128), ("encoder_bottleneck", 128, 64)]
    def dec(p):
        return [(f"{p}_d1", 64, 128), (f"{p}_d2", 128, 256), (f"{p}_d3", 256, 512),
                (f"{p}_reconstruction", 512, c)]
    return enc + dec("main") + dec("reg")


def init_params(num_cards: int, seed: int = 0, dtype=np.float32):
    """glorot_uniform kernels / zero biases from ``torch.Generator().manual_seed(seed)``
    (SURVEY.md §8d); returns ``{name+'/kernel': (in,out), name+'/bias': (out,)}``."""
    import torch
    from cubecobrarecommender_b200.synth import glorot_uniform
    gen = torch.Generator().manual_seed(seed)
    params = {}
    for name, fi, fo in layer_specs(num_cards):
        params[name + "/kernel"] = glorot_uniform(fi, fo, gen).astype(dtype)
        params[name + "/bias"] = np.zeros(fo, dtype=dtype)
    return params


# ------------------------------------------------------------ NumPy float64
def _dense_fwd(x, w, b, relu=True):
    z = x @ w + b
    return np.maximum(z, 0) if relu else z


def forward_np(params, x, reg_rows, onehot_rows_as_gather=False):
    """Returns (logits_main (B,C), logits_reg (R,C), cache).  ``x`` dense
    (B,C); ``reg_rows`` int (R,) = the rows of I fed to the second tower
    (reference ``model.py:117-125``; ``generator.py:47-51,76``).

    ``onehot_rows_as_gather``: evaluate the first layer of the second tower as
    ``relu(W1[r] + b1)`` instead of ``relu(I[r] @ W1 + b1)`` -- the same float64
    numbers (a one-hot row contributes one exact product and exact zeros) without
    the (R, C) one-hot matrix; for the BASELINE-sized parity cases."""
    p = {k: v.astype(np.float64) for k, v in params.items()}
    c = p["encoder_e1/kernel"].shape[0]
    reg_rows = np.asarray(reg_rows, dtype=np.int64)
    if onehot_rows_as_gather:
        eye_rows = None
    else:
        eye_rows = np.zeros((len(reg_rows), c)); eye_rows[np.arange(len(reg_rows)), reg_rows] = 1
    cache = {}
    outs = []
    for tower, inp, dec in (("main", np.asarray(x, np.float64), "main"), ("reg", eye_rows, "reg")):
        acts = [inp]
        names = ["encoder_e1", "encoder_e2", "encoder_e3", "encoder_bottleneck",
                 f"{dec}_d1", f"{dec}_d2", f"{dec}_d3"]
        for n in names:
            if acts[-1] is None:        # one-hot rows through the first layer = a row gather of its kernel
                acts.append(np.maximum(p[n + "/kernel"][reg_rows] + p[n + "/bias"], 0))
            else:
                acts.append(_dense_fwd(acts[-1], p[n + "/kernel"], p[n + "/bias"]))
        z = _dense_fwd(acts[-1], p[f"{dec}_reconstruction/kernel"],
                       p[f"{dec}_reconstruction/bias"], relu=False)
        cache[tower] = (names + [f"{dec}_reconstruction"], acts)
        outs.append(z)
    return outs[0], outs[1], cache


def bce_from_logits_np(z, y):
    """mean_{b,c} of max(z,0) - z*y + log1p(exp(-|z|))  [Keras-2.5 logits path]."""
    return float(np.mean(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))))


def softmax_np(z):
    z = z - z.max(1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(1, keepdims=True)


def kld_np(t, q):
    """mean_r sum_c t' log(t'/q'), t',q' clipped to [1e-7, 1]  [Keras-2.5]."""
    tc = np.clip(t, KERAS_EPS, 1.0)
    qc = np.clip(q, KERAS_EPS, 1.0)
    return float(np.mean(np.sum(tc * np.log(tc / qc), axis=1)))


def loss_and_grads_np(params, x, y, reg_rows, t_reg, reg, onehot_rows_as_gather=False):
    """float64 loss (total, bce, kl) and hand-derived gradients for every
    parameter.  ``t_reg`` = M-hat[reg_rows] (R,C).  Follows SURVEY.md §8a-6.
    ``onehot_rows_as_gather``: see :func:`forward_np` (the first-layer gradient of
    the one-hot rows is then a row scatter-add, again the same numbers)."""
    z1, z2, cache = forward_np(params, x, reg_rows, onehot_rows_as_gather)
    y = np.asarray(y, np.float64); t = np.asarray(t_reg, np.float64)
    b, c = z1.shape; r = z2.shape[0]
    bce = bce_from_logits_np(z1, y)
    q = softmax_np(z2)
    kl = kld_np(t, q)
    total = bce + reg * kl
    # dL/dz1 = (sigmoid(z) - y) / (B*C)
    dz1 = (1.0 / (1.0 + np.exp(-z1)) - y) / (b * c)
    # dL/dz2 through clip + softmax: (q*S - t'*1[unclipped]) / R, S = sum_unclipped t'
    tc = np.clip(t, KERAS_EPS, 1.0)
    unclipped = (q >= KERAS_EPS) & (q <= 1.0)
    s = (tc * unclipped).sum(1, keepdims=True)
    dz2 = reg * (q * s - tc * unclipped) / r
    grads = {k: np.zeros_like(v, dtype=np.float64) for k, v in params.items()}
    p = {k: v.astype(np.float64) for k, v in params.items()}
    for tower, dz in (("main", dz1), ("reg", dz2)):
        names, acts = cache[tower]
        d = dz
        for li in range(len(names) - 1, -1, -1):
            n = names[li]
            a_in = acts[li]
            if a_in is None:            # one-hot input rows: I[r]^T d adds row i of d to row r_i of the kernel gradient
                np.add.at(grads[n + "/kernel"], np.asarray(reg_rows, dtype=np.int64), d)
            else:
                grads[n + "/kernel"] += a_in.T @ d
            grads[n + "/bias"] += d.sum(0)
            if li > 0:
                d = d @ p[n + "/kernel"].T
                d = d * (acts[li] > 0)      # acts[li] is the ReLU output feeding layer li
    return (total, bce, kl), grads


def binary_accuracy_np(z1, y):
    """Keras ``metrics=['accuracy']`` on the sigmoid output trained with binary_crossentropy (reference
    ``train.py:87``) resolves to ``binary_accuracy``: mean over all cells of ``(sigmoid(z) > 0.5) == y``  [Keras-2.5
    metrics.binary_accuracy, threshold 0.5], i.e. ``(z > 0) == y``."""
    return float(np.mean((np.asarray(z1) > 0) == (np.asarray(y) != 0)))


def categorical_accuracy_np(z2, t):
    """The same list entry on the softmax output (target = a probability row) resolves to ``categorical_accuracy``:
    mean over rows of ``argmax(y_true) == argmax(y_pred)``, first maximal index on ties  [Keras-2.5]."""
    return float(np.mean(np.argmax(np.asarray(t), axis=1) == np.argmax(np.asarray(z2), axis=1)))


def adam_step_np(params, grads, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=KERAS_EPS):
    """TF/Keras-2.5 ``Adam`` dense update at (1-based) step ``t``; in place."""
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for k in params:
        g = grads[k].astype(params[k].dtype)
        m[k] = b1 * m[k] + (1 - b1) * g
        v[k] = b2 * v[k] + (1 - b2) * g * g
        params[k] = params[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
    return params, m, v


def recommend_np(params, x):
    """``decoder(encoder(x))`` = sigmoid probabilities (reference
    ``ml_recommend.py:78-80``), float64."""
    z1, _, _ = forward_np(params, x, np.zeros(0, dtype=np.int64))
    return 1.0 / (1.0 + np.exp(-z1))


def rank_additions(results, in_cube, amount):
    """Reference ``ml_recommend.py:87-104``: ``argsort()[::-1]``, skip in-cube,
    first ``amount``.  Tie rule: stable argsort reversed (descending score,
    ties -> larger index first)."""
    ranked = np.asarray(results).argsort(kind="stable")[::-1]
    out = []
    for rec in ranked:
        if in_cube[rec] != 1:
            out.append(int(rec))
            if len(out) >= amount:
                break
    return out


def card_embeddings_np(params):
    """``model.encoder(I)`` (reference ``src/scripts/similarity.py:19-24``): the encoder applied to every
    one-hot card, i.e. ``relu(W1 + b1)`` pushed through the three remaining encoder layers.  float64 (C, 64)."""
    p = {k: v.astype(np.float64) for k, v in params.items()}
    h = np.maximum(p["encoder_e1/kernel"] + p["encoder_e1/bias"], 0.0)
    for n in ("encoder_e2", "encoder_e3", "encoder_bottleneck"):
        h = np.maximum(h @ p[n + "/kernel"] + p[n + "/bias"], 0.0)
    return h


def similarity_np(embs, idx):
    """Keras ``CosineSimilarity()(embs[idx], x)`` for every row x (reference ``similarity.py:27-29``):
    ``-sum(l2_normalize(a) * l2_normalize(b))`` with ``l2_normalize(v) = v / sqrt(max(sum v^2, 1e-12))``
    [Keras-2.5 losses.cosine_similarity].  The caller ranks with ``argsort()`` ascending."""
    n = embs / np.sqrt(np.maximum((embs * embs).sum(1, keepdims=True), 1e-12))
    return -(n @ n[idx])


# ---------------------------------------------------------------- torch CPU
class TorchDAE:
    """torch restatement used (a) as the second, autograd formulation and
    (b) as the CPU comparator for the train step (all host threads)."""

    def __init__(self, params, dtype=None, device="cpu"):
        import torch
        self.torch = torch
        dtype = dtype or torch.float32
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype, device=device).requires_grad_(True)
                  for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.t = 0

    def tower(self, inp, dec):
        torch = self.torch
        h = inp
        for n in ("encoder_e1", "encoder_e2", "encoder_e3", "encoder_bottleneck",
                  f"{dec}_d1", f"{dec}_d2", f"{dec}_d3"):
            h = torch.relu(h @ self.p[n + "/kernel"] + self.p[n + "/bias"])
        return h @ self.p[f"{dec}_reconstruction/kernel"] + self.p[f"{dec}_reconstruction/bias"]

    def loss(self, x, y, reg_rows, t_reg, reg):
        torch = self.torch
        z1 = self.tower(x, "main")
        # one-hot rows of I through the dense first layer == rows of the kernel
        w1 = self.p["encoder_e1/kernel"]
        eye_rows = torch.zeros(len(reg_rows), w1.shape[0], dtype=w1.dtype, device=w1.device)
        eye_rows[torch.arange(len(reg_rows)), reg_rows] = 1
        z2 = self.tower(eye_rows, "reg")
        bce = (torch.clamp(z1, min=0) - z1 * y + torch.log1p(torch.exp(-z1.abs()))).mean()
        q = torch.softmax(z2, dim=1)
        tc = torch.clamp(t_reg, KERAS_EPS, 1.0)
        qc = torch.clamp(q, KERAS_EPS, 1.0)
        kl = (tc * torch.log(tc / qc)).sum(1).mean()
        return bce + reg * kl, bce, kl

    def train_step(self, x, y, reg_rows, t_reg, reg, lr=1e-3, b1=0.9, b2=0.999, eps=KERAS_EPS):
        torch = self.torch
        for v in self.p.values():
            v.grad = None
        total, bce, kl = self.loss(x, y, reg_rows, t_reg, reg)
        total.backward()
        self.t += 1
        lr_t = lr * (1.0 - b2 ** self.t) ** 0.5 / (1.0 - b1 ** self.t)
        with torch.no_grad():
            for k, w in self.p.items():
                g = w.grad
                self.m[k].mul_(b1).add_(g, alpha=1 - b1)
                self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                w.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + eps))
        return float(total), float(bce), float(kl)

    def grads(self):
        return {k: v.grad.detach().cpu().numpy() for k, v in self.p.items()}

    def params_np(self):
        return {k: v.detach().cpu().numpy() for k, v in self.p.items()}
