"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

What is executed from the reference:
  * ``utils.create_adjacency_matrix``  (src/non_ml/utils.py:75-92)  -- plain import
  * ``simple_recs`` / ``simple_cuts``  (src/scripts/recommend.py:7-18,
    src/scripts/cut_cards.py:7-18)     -- the FunctionDef is extracted with ``ast``
    because the scripts fetch a URL at import time
  * ``DataGenerator``                  (src/ml/generator.py:4-103) -- imported behind
    a stub ``tensorflow.keras.utils.Sequence`` (its only TF dependency, line 1)
The Keras model/train code cannot run here (no TensorFlow), so no golden exists
for it ("parity unpinned" for that half; see oracle/dae.py).
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


def _extract_function(path, name):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def _import_reference_generator():
    tf = types.ModuleType("tensorflow")
    keras = types.ModuleType("tensorflow.keras")
    kutils = types.ModuleType("tensorflow.keras.utils")

    class Sequence:  # the base class is the only thing generator.py needs
        pass

    kutils.Sequence = Sequence
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras,
                        "tensorflow.keras.utils": kutils})
    sys.path.insert(0, os.path.join(REF, "src", "ml"))
    import generator  # noqa
    return generator.DataGenerator


def small_cubes(k, c, seed, lo, hi):
    from cubecobrarecommender_b200.synth import synth_cubes_csr, csr_to_dense
    indptr, indices = synth_cubes_csr(k, c, size_lo=lo, size_hi=hi, seed=seed)
    return indptr, indices, csr_to_dense(indptr, indices, c)


def main():
    sys.path.insert(0, os.path.join(REF, "src", "non_ml"))
    import utils as ref_utils
    simple_recs = _extract_function(os.path.join(REF, "src/scripts/recommend.py"), "simple_recs")
    simple_cuts = _extract_function(os.path.join(REF, "src/scripts/cut_cards.py"), "simple_cuts")
    RefGen = _import_reference_generator()

    # ---- graph: 80 cubes x 131 cards, cards 120..130 never drawn (zero rows) ----
    k, c_live, c = 80, 120, 131
    indptr, indices, dense_live = small_cubes(k, c_live, seed=1234, lo=8, hi=40)
    dense = np.zeros((k, c)); dense[:, :c_live] = dense_live
    adj = ref_utils.create_adjacency_matrix(dense, verbose=False)
    adj_fd = ref_utils.create_adjacency_matrix(dense, verbose=False, force_diag=0.5)
    rec_cubes = dense[[0, 7, 33]]
    recs = np.stack([np.array(simple_recs(cu, adj), dtype=np.int64)[:50] for cu in rec_cubes])
    # the float64 scores simple_recs ranks, evaluated with the reference's own
    # expression (recommend.py:8-13) so NumPy's pairwise summation order is kept
    def _scores(cu):
        contains = np.where(cu == 1)[0]; missing = np.where(cu == 0)[0]
        out = np.full(c, -np.inf)
        out[missing] = adj[contains][:, missing].sum(0)
        return out
    rec_scores = np.stack([_scores(cu) for cu in rec_cubes])
    cuts = [np.array(simple_cuts(cu, adj.copy()), dtype=np.int64) for cu in rec_cubes]
    cuts_pad = np.full((3, max(len(x) for x in cuts)), -1, dtype=np.int64)
    for i, x in enumerate(cuts):
        cuts_pad[i, :len(x)] = x
    np.savez_compressed(
        os.path.join(HERE, "graph_small.npz"),
        indptr=indptr, indices=indices, num_cards=np.int64(c),
        adj=adj, adj_force_diag=adj_fd, rec_cube_rows=np.array([0, 7, 33]),
        recs_top50=recs, rec_scores=rec_scores, cuts=cuts_pad)

    # ---- graph, larger cubes (130..300 of 400 cards): exercises NumPy's recursive
    # pairwise summation (n > 128) inside simple_recs / simple_cuts ----
    k2, c2 = 60, 400
    ip2, ix2, dense2 = small_cubes(k2, c2, seed=99, lo=130, hi=300)
    adj2 = ref_utils.create_adjacency_matrix(dense2, verbose=False)
    rows2 = np.array([3, 41])
    recs2 = np.stack([np.array(simple_recs(dense2[r], adj2), dtype=np.int64)[:50] for r in rows2])
    def _scores2(cu):
        contains = np.where(cu == 1)[0]; missing = np.where(cu == 0)[0]
        out = np.full(c2, -np.inf)
        out[missing] = adj2[contains][:, missing].sum(0)
        return out
    def _cutscores2(cu):
        a0 = adj2.copy(); np.fill_diagonal(a0, 0)
        contains = np.where(cu == 1)[0]
        out = np.full(c2, np.inf)
        out[contains] = a0[contains][:, contains].sum(0)
        return out
    cuts2 = [np.array(simple_cuts(dense2[r], adj2.copy()), dtype=np.int64) for r in rows2]
    cuts2_pad = np.full((2, max(len(x) for x in cuts2)), -1, dtype=np.int64)
    for i, x in enumerate(cuts2):
        cuts2_pad[i, :len(x)] = x
    np.savez_compressed(
        os.path.join(HERE, "graph_pairwise.npz"),
        indptr=ip2, indices=ix2, num_cards=np.int64(c2), rec_cube_rows=rows2,
        recs_top50=recs2, rec_scores=np.stack([_scores2(dense2[r]) for r in rows2]),
        cuts=cuts2_pad, cut_scores=np.stack([_cutscores2(dense2[r]) for r in rows2]))

    # ---- noise: reference DataGenerator, seeded global MT19937 ----
    y_mtx = adj.copy(); np.fill_diagonal(y_mtx, 1); y_mtx = y_mtx / y_mtx.sum(1)[:, None]  # train.py:69-71
    np.random.seed(4242)
    gen = RefGen(y_mtx, dense, batch_size=16, noise=0.2)
    (x0, xr0), (y0, yr0) = gen[0]
    (x1, xr1), (y1, yr1) = gen[2]
    np.savez_compressed(
        os.path.join(HERE, "noise_small.npz"),
        seed=np.int64(4242), batch_size=np.int64(16), y_mtx=y_mtx,
        neg_sampler=gen.neg_sampler, epoch_indices=gen.indices,
        x0=x0.astype(np.int8), y0=y0.astype(np.int8), reg0=np.argmax(xr0, 1), yr0=yr0,
        x1=x1.astype(np.int8), y1=y1.astype(np.int8), reg1=np.argmax(xr1, 1), yr1=yr1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
