"""The C-ABI library builds, loads and exports every symbol include/cubecobra_b200.h
declares (no compute calls: this runs without a GPU)."""
import os
import re

import numpy as np
import pytest

from cubecobrarecommender_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(REPO, "include", "cubecobra_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = _declared()
    assert declared, "no declarations parsed"
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_version_and_error_channel(lib):
    assert lib.cc_version() == 100
    # argument errors are reported without touching the GPU
    rc = lib.cc_alias_build_host(None, 4, None, None)
    assert rc == -1 and "cc_alias_build_host" in _lib.last_error()


def test_size_helpers(lib):
    assert lib.cc_bits_words(20000) == 640 and lib.cc_bits_words(0) == 0
    assert lib.cc_bits_cpad(21000) == 21120
    # NumPy pairwise leaf plan: n <= 128 is one leaf; 540 rows split into 7 leaves
    assert lib.cc_pairwise_leaf_count(128) == 1 and lib.cc_pairwise_leaf_count(129) == 2
    ptr_h = np.array([0, 540], dtype=np.int64)
    n = lib.cc_pairwise_leaf_count(540)
    plan = np.zeros((n, 4), dtype=np.int32); leaf_ptr = np.zeros(2, dtype=np.int32)
    _lib.call("cc_pairwise_plan_host", _lib.ptr(ptr_h), 1, _lib.ptr(plan), _lib.ptr(leaf_ptr))
    assert leaf_ptr[1] == n and plan[:, 1].sum() == 540 and (plan[:, 1] <= 128).all()
    assert plan[:, 2].sum() == n - 1          # a binary tree over n leaves has n-1 merges
    assert (plan[1:, 0] == np.cumsum(plan[:-1, 1])).all()


def test_alias_table_is_exact(lib):
    rng = np.random.default_rng(0)
    p = rng.random(257); p[5] = 0.0; p /= p.sum()
    prob = np.zeros(257, np.float32); alias = np.zeros(257, np.int32)
    _lib.call("cc_alias_build_host", _lib.ptr(p), 257, _lib.ptr(prob), _lib.ptr(alias))
    mass = prob.astype(np.float64) / 257
    np.add.at(mass, alias, (1.0 - prob.astype(np.float64)) / 257)
    assert np.abs(mass - p).max() < 1e-7


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.CubeCobraError):
        _lib.load(str(tmp_path / "nope.so"))
