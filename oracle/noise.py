"""Oracle (TEST INFRASTRUCTURE): the noise function F and batch assembly.

CPU restatement of ``DataGenerator`` -- reference ``src/ml/generator.py:4-103``.
It draws from the *global* ``numpy.random`` MT19937 state with the same calls in
the same order as the reference, so after ``np.random.seed(s)`` it reproduces
the reference's batches bit for bit (pinned by ``tests/golden/noise_small.npz``).

The CUDA noise kernel uses Philox streams and cannot reproduce MT19937 draws;
against it this oracle supplies the *distribution* and the invariants
(SURVEY.md §8a-3), while train-step parity is checked with the noise output
held fixed.
"""
from __future__ import annotations

import numpy as np


class DataGenerator:
    """Reference ``generator.py:6-31`` (constructor), same argument order."""

    def __init__(self, adj_mtx, cubes, batch_size=64, shuffle=True, to_fit=True,
                 noise=0.2, noise_std=0.1):
        self.noise_std = noise_std
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.to_fit = to_fit
        self.noise = noise
        self.y_reg = adj_mtx                                   # generator.py:22
        self.x_reg = np.zeros_like(adj_mtx)                    # generator.py:23-24
        np.fill_diagonal(self.x_reg, 1)
        self.x_main = cubes
        self.N_cubes = self.x_main.shape[0]
        self.N_cards = self.x_main.shape[1]
        self.reset_indices()                                   # generator.py:29
        self.neg_sampler = adj_mtx.sum(0) / adj_mtx.sum()      # generator.py:30

    def __len__(self):                                         # generator.py:32-36
        return self.N_cubes // self.batch_size

    def reset_indices(self):                                   # generator.py:63-66
        self.indices = np.arange(self.N_cubes)
        if self.shuffle == True:  # noqa: E712  (as written in the reference)
            np.random.shuffle(self.indices)

    def on_epoch_end(self):                                    # generator.py:68-72
        self.reset_indices()

    def __getitem__(self, batch_number):                       # generator.py:38-61
        main_indices = self.indices[
            batch_number * self.batch_size:(batch_number + 1) * self.batch_size]
        reg_indices = np.random.choice(np.arange(self.N_cards), len(main_indices),
                                       p=self.neg_sampler)
        x, y = self.generate_data(main_indices, reg_indices)
        if self.to_fit:
            return [x[0], x[1]], [y[0], y[1]]
        return [x[0], x[1]]

    def generate_data(self, main_indices, reg_indices):        # generator.py:74-103
        cubes = self.x_main[main_indices]
        x_regularization = self.x_reg[reg_indices]
        y_regularization = self.y_reg[reg_indices]
        cut_mask = np.zeros((self.batch_size, self.N_cards))
        add_mask = np.zeros((self.batch_size, self.N_cards))
        y_cut_mask = np.zeros((self.batch_size, self.N_cards))
        for i, cube in enumerate(cubes):
            includes = np.where(cube == 1)[0]
            excludes = np.where(cube == 0)[0]
            size = len(includes)
            noise = np.clip(np.random.normal(self.noise, self.noise_std),
                            a_min=0.05, a_max=0.8)
            flip_amount = int(size * noise)
            flip_include = np.random.choice(includes, flip_amount)
            p = self.neg_sampler[excludes] / self.neg_sampler[excludes].sum()
            flip_exclude = np.random.choice(excludes, flip_amount, p=p)
            y_flip_include = np.random.choice(flip_include, flip_amount // 4)
            cut_mask[i, flip_include] = -1
            y_cut_mask[i, y_flip_include] = -1
            add_mask[i, flip_exclude] = 1
        x_cubes = cubes + cut_mask + add_mask
        y_cubes = cubes + y_cut_mask
        return [(x_cubes, x_regularization), (y_cubes, y_regularization)]


def check_noise_invariants(cubes, x_cubes, y_cubes):
    """The invariants of F: x, y in {0,1}; cards removed
    from y are a subset of the cards removed from x; added cards were not in
    the cube; y never adds.  Raises AssertionError otherwise."""
    cubes = np.asarray(cubes); x = np.asarray(x_cubes); y = np.asarray(y_cubes)
    assert np.isin(x, (0, 1)).all() and np.isin(y, (0, 1)).all()
    removed_x = (cubes == 1) & (x == 0)
    removed_y = (cubes == 1) & (y == 0)
    added_x = (cubes == 0) & (x == 1)
    assert not (removed_y & ~removed_x).any(), "y removed a card x kept"
    assert not ((cubes == 0) & (y == 1)).any(), "y added a card"
    # flips: distinct removed <= flip, distinct added <= flip, flip <= 0.8*size
    size = (cubes == 1).sum(1)
    assert (removed_x.sum(1) <= np.floor(0.8 * size)).all()
    assert (added_x.sum(1) <= np.floor(0.8 * size)).all()
    return removed_x, removed_y, added_x
