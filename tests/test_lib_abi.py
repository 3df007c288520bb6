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


def test_gemm_planner_choices(lib):
    """Host-side planner of the tcgen05 GEMMs (no GPU needed; 148 SMs assumed without a device): the 512 <-> C
    passes of the train step run on 256 x 256 CTA-pair tiles with enough tiles to fill the 74 pairs, the small layers
    stay on single CTAs, explicit requests are honoured."""
    def plan(m, n, k, precision=1, tile_n=0, split_k=0):
        out = np.zeros(3, dtype=np.int32)
        _lib.call("cc_gemm_tc_plan", precision, m, n, k, tile_n, split_k, _lib.ptr(out))
        return tuple(int(v) for v in out)
    B, C, H = 4096, 20884, 512
    for m, n, k in ((B, C, H), (H, C, B), (B, H, C), (C, H, B)):          # fwd, dW, dX, dW1
        bn, split, ctas = plan(m, n, k)
        assert (bn, ctas) == (256, 2)
        tiles = -(-m // 256) * -(-n // 256) * split
        sk = np.zeros(4, dtype=np.int32)
        _lib.call("cc_gemm_tc_plan_ex", 1, m, n, k, 0, 0, _lib.ptr(sk))
        assert tiles >= 74 or sk[3] == 1       # a full wave of CTA pairs, or stream-K spans spread over all of them
    def plan4(m, n, k, precision=1):
        out = np.zeros(4, dtype=np.int32)
        _lib.call("cc_gemm_tc_plan_ex", precision, m, n, k, 0, 0, _lib.ptr(out))
        return tuple(int(v) for v in out)
    # the forward pass (16 k-blocks, bias epilogue) stays data-parallel and unsplit; the three gradient passes (K = 4096
    # or 20 884: ragged tile waves) take the hybrid stream-K schedule, which comes with K split 1
    assert plan4(B, C, H) == (256, 1, 2, 0)
    for m, n, k in ((H, C, B), (B, H, C), (C, H, B)):
        assert plan4(m, n, k) == (256, 1, 2, 1) and plan4(m, n, k, precision=2) == (256, 1, 2, 1)
    _lib.call("cc_gemm_tc_set_stream_k", 0)
    try:
        assert plan(B, C, H)[1] == 1 and plan(B, H, C)[1] > 1 and plan4(B, H, C)[3] == 0      # plain split-K when switched off
    finally:
        _lib.call("cc_gemm_tc_set_stream_k", -1)
    for m, n, k in ((8192, 256, 512), (8192, 128, 256), (4096, 512, 256), (512, 256, 8192), (64, 128, 4096)):
        assert plan(m, n, k)[2] == 1                                        # small layers: latency-bound, single CTAs
    assert plan(8192, 64, 128)[0] == 128                                    # narrow outputs never take 256-wide tiles
    assert plan(B, C, H, tile_n=128, split_k=1) == (128, 1, 1)
    assert plan(B, H, C, tile_n=256, split_k=7)[:2] == (256, 7)
    assert plan(B, C, H, precision=2)[0] == 256


def test_chain_and_readable_range_argument_checks(lib):
    """Host-side argument checking of the round-2 entry points (no device work is reached): cc_chain_tc rejects layer
    counts / widths it cannot run, cc_gemm_tc_register_readable takes and forgets ranges."""
    import ctypes
    widths = np.array([512, 256, 128, 64], dtype=np.int32)
    dummy = np.zeros(3, dtype=np.uint64)
    ld = np.zeros(3, dtype=np.int64)
    kn = np.zeros(3, dtype=np.int32)
    args = (_lib.ptr(dummy), _lib.ptr(ld), _lib.ptr(kn), None, None, None, None, None, 1, _lib.ptr(dummy), _lib.ptr(ld), 1, None)
    rc = lib.cc_chain_tc(128, 0, _lib.ptr(widths), ctypes.c_void_p(16), 512, *args)
    assert rc == -1 and "1..3 layers" in _lib.last_error()
    bad = np.array([500, 256], dtype=np.int32)                 # k must be a multiple of 32
    rc = lib.cc_chain_tc(128, 1, _lib.ptr(bad), ctypes.c_void_p(16), 512, *args)
    assert rc == -1 and "multiple of 32" in _lib.last_error()
    wide = np.array([64, 512, 128], dtype=np.int32)            # only the last layer may be 512 wide
    rc = lib.cc_chain_tc(128, 2, _lib.ptr(wide), ctypes.c_void_p(16), 64, *args)
    assert rc == -1 and "last layer" in _lib.last_error()
    assert lib.cc_gemm_tc_register_readable(None, 64) == -1
    assert lib.cc_gemm_tc_register_readable(ctypes.c_void_p(4096), 1 << 20) == 0
    assert lib.cc_gemm_tc_register_readable(ctypes.c_void_p(4096), 0) == 0      # forgotten again
    assert lib.cc_gemm_tc_mn3_count() >= 0
