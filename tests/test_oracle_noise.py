"""Oracle DataGenerator vs the reference's (tests/golden/noise_small.npz), bit for bit
(reference src/ml/generator.py:38-103)."""
import numpy as np

from cubecobrarecommender_b200.synth import csr_to_dense
from oracle import noise as onoise


def _dense(graph_golden):
    g = graph_golden
    c = int(g["num_cards"])
    dense = np.zeros((len(g["indptr"]) - 1, c))
    dense[:, :120] = csr_to_dense(g["indptr"], g["indices"], 120)
    return dense


def test_generator_reproduces_reference_batches(graph_golden, noise_golden):
    n = noise_golden
    dense = _dense(graph_golden)
    np.random.seed(int(n["seed"]))
    gen = onoise.DataGenerator(n["y_mtx"], dense, batch_size=int(n["batch_size"]), noise=0.2)
    assert np.array_equal(gen.indices, n["epoch_indices"])
    assert np.array_equal(gen.neg_sampler, n["neg_sampler"])
    assert len(gen) == 80 // 16
    (x0, xr0), (y0, yr0) = gen[0]
    (x1, xr1), (y1, yr1) = gen[2]
    assert np.array_equal(x0, n["x0"]) and np.array_equal(y0, n["y0"])
    assert np.array_equal(x1, n["x1"]) and np.array_equal(y1, n["y1"])
    assert np.array_equal(np.argmax(xr0, 1), n["reg0"]) and np.array_equal(yr0, n["yr0"])
    assert np.array_equal(np.argmax(xr1, 1), n["reg1"]) and np.array_equal(yr1, n["yr1"])


def test_noise_invariants_on_golden(graph_golden, noise_golden):
    n = noise_golden
    dense = _dense(graph_golden)
    cubes0 = dense[n["epoch_indices"][:16]]
    onoise.check_noise_invariants(cubes0, n["x0"], n["y0"])
    cubes1 = dense[n["epoch_indices"][32:48]]
    onoise.check_noise_invariants(cubes1, n["x1"], n["y1"])
