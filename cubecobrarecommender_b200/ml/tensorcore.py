"""tcgen05 tensor-core GEMM wrappers (precision modes "tf32" and "bf16"); see csrc/gemm_tc.cu."""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import call, ptr, stream_ptr

PRECISION_CODE = {"tf32": 1, "bf16": 2}
_SMS = None


def _sms():
    global _SMS
    if _SMS is None:
        _SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _SMS


def gemm(a, b, c, *, transa=False, transb=False, bias=None, relu=False, mask=None, accumulate=False,
         precision="tf32", split_k=0, tile_n=0, round_out=False):
    """split_k = 0 / tile_n = 0: the library picks the tile width (128 | 256) and the K split from a
    wave-quantisation model (csrc/gemm_tc.cu plan_eff).  ``round_out``: round the fp32 output to tf32 (it feeds a
    kind::tf32 GEMM), whatever the operand kind of this GEMM."""
    m, n = c.shape
    k = a.shape[0] if transa else a.shape[1]
    want = torch.float32 if precision == "tf32" else torch.bfloat16
    if a.dtype != want or b.dtype != want:
        raise TypeError(f"precision {precision} needs {want} operands, got {a.dtype} / {b.dtype}")
    call("cc_gemm_tc", PRECISION_CODE[precision], int(transa), int(transb), m, n, k, ptr(a), a.stride(0), ptr(b),
         b.stride(0), ptr(c), c.stride(0), ptr(bias), int(relu), ptr(mask), mask.stride(0) if mask is not None else 0,
         int(accumulate), int(split_k or 0), int(tile_n), int(round_out), stream_ptr())
    return c


def gemm_bce(a, w, bias, ybits, count, dz, loss_partial, precision="tf32", round_out=True, dbias=None, acc_partial=None):
    """Fused  z = a @ w + bias -> BCE loss partials + dlogits  (z never stored); ``dbias`` (optional, float [n])
    receives the column sums of dlogits = the gradient of ``bias``; ``acc_partial`` (optional, float64 like
    ``loss_partial``) the counts of correctly rounded cells (Keras binary_accuracy)."""
    m, k = a.shape
    n = w.shape[1]
    dz16 = dz.dtype == torch.bfloat16        # bf16 dlogits for the bf16 dW / dX GEMMs ("bf16" mode)
    if dz16 and precision != "bf16":
        raise TypeError("bf16 dlogits come with bf16 operands")
    call("cc_gemm_bce_tc_ex", PRECISION_CODE[precision], m, n, k, ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(bias),
         ptr(ybits), ybits.stride(0), float(count), ptr(dz), dz.stride(0), ptr(loss_partial), ptr(dbias),
         int(round_out and precision == "tf32"), int(dz16), ptr(acc_partial), stream_ptr())


def bce_partial_count(m, lddz):
    return _lib.load().cc_gemm_bce_partial_count(m, lddz)
