"""Regularised denoising auto-encoder: host-side mirror of reference ``src/ml/model.py``.

``CC_Recommender(num_cards)`` keeps the reference's surface -- attributes ``encoder``,
``decoder``, ``decoder_for_reg`` and ``call((x, identity)) -> (reconstruction,
decoded_for_reg)`` (reference ``model.py:89-125``) -- over flat float32 parameter buffers in
HBM and the kernels of libcubecobra_b200.so.  The 24 tensors carry the Keras names
(``encoder_e1/kernel`` ... ``reg_reconstruction/bias``), kernels stored ``(in, out)``.

``DAEEngine`` is the train step of reference ``src/ml/train.py:83-102`` (one ``fit`` step):
forward of both towers, BCE + reg*KLD, backward, TF-style Adam.
"""
from __future__ import annotations


import numpy as np
import torch

from .. import _lib
from .._lib import call, ptr, stream_ptr
from ..sparse import CubeCSR

HIDDEN = (512, 256, 128, 64)
ENC_NAMES = ("encoder_e1", "encoder_e2", "encoder_e3", "encoder_bottleneck")


def dec_names(prefix):
    return (f"{prefix}_d1", f"{prefix}_d2", f"{prefix}_d3", f"{prefix}_reconstruction")


def layer_specs(num_cards: int):
    c = num_cards
    enc = [("encoder_e1", c, 512), ("encoder_e2", 512, 256), ("encoder_e3", 256, 128),
           ("encoder_bottleneck", 128, 64)]
    def dec(p):
        return [(f"{p}_d1", 64, 128), (f"{p}_d2", 128, 256), (f"{p}_d3", 256, 512),
                (f"{p}_reconstruction", 512, c)]
    return enc + dec("main") + dec("reg")


def _round4(n):
    return (n + 3) // 4 * 4


class ParamStore:
    """Flat float32 buffers (params, grads, Adam m and v) with named 2-D/1-D views.
    One buffer = one Adam launch and one NCCL all_reduce."""

    def __init__(self, num_cards: int, device):
        self.num_cards = num_cards
        self.device = torch.device(device)
        self.layout = {}
        off = 0
        for name, fi, fo in layer_specs(num_cards):
            self.layout[name + "/kernel"] = (off, (fi, fo)); off += _round4(fi * fo)
            self.layout[name + "/bias"] = (off, (fo,)); off += _round4(fo)
        self.total = off
        self.params = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.adam_m = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.adam_v = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.step = torch.zeros(1, dtype=torch.int64, device=self.device)   # completed steps
        self.shadow = None        # tf32 (round-to-nearest) copy of params for the tensor-core GEMMs

    def enable_tf32_shadow(self):
        if self.shadow is None:
            self.shadow = torch.zeros_like(self.params)
            # the GEMMs read Keras kernels out of this buffer; a kernel's last row is followed by its bias, so the single-TMA
            # form of an MN-major operand may read the few floats behind a row that is not a multiple of 32 wide
            call("cc_gemm_tc_register_readable", ptr(self.shadow), self.shadow.numel() * 4)
        self.sync_shadow()

    def __del__(self):
        try:
            if getattr(self, "shadow", None) is not None:
                call("cc_gemm_tc_register_readable", ptr(self.shadow), 0)
        except Exception:
            pass

    def sync_shadow(self):
        if self.shadow is not None:
            call("cc_round_tf32", ptr(self.params), ptr(self.shadow), self.total, stream_ptr())

    # -- data parallel over NVLink peer memory ---------------------------------------------------
    def make_symmetric(self, group=None):
        """Move params and grads into symmetric memory (torch.distributed._symmetric_memory) so that every rank
        can load any rank's gradients and store into any rank's parameters from inside a kernel
        (cc_adam_step_p2p).  Collective: every rank of ``group`` must call it."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        import warnings
        with warnings.catch_warnings():             # deprecated no-op on newer torch, required on older ones
            warnings.simplefilter("ignore")
            try:
                symm.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
        new_p = symm.empty(self.total, dtype=torch.float32, device=self.device)
        new_g = symm.empty(self.total, dtype=torch.float32, device=self.device)
        new_p.copy_(self.params)
        new_g.zero_()
        self._p_hdl = symm.rendezvous(new_p, group)
        self._g_hdl = symm.rendezvous(new_g, group)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        pp = [int(x) for x in self._p_hdl.buffer_ptrs]
        gp = [int(x) for x in self._g_hdl.buffer_ptrs]
        if pp[rank] != new_p.data_ptr() or gp[rank] != new_g.data_ptr() or len(pp) != world:
            raise RuntimeError("symmetric-memory rendezvous returned unexpected buffer pointers")
        self.params, self.grads = new_p, new_g
        # NVLS multicast mappings of the two buffers (0 when the fabric has no multicast support)
        self.mc_params = int(getattr(self._p_hdl, "multicast_ptr", 0) or 0)
        self.mc_grads = int(getattr(self._g_hdl, "multicast_ptr", 0) or 0)
        self.peer_params = np.array(pp, dtype=np.uint64)
        self.peer_grads = np.array(gp, dtype=np.uint64)
        self.dp_rank, self.dp_world = rank, world
        from ..dist import owner_slice
        self.dp_slice = owner_slice(self.total, rank, world)   # 16-byte aligned (offsets are padded to 4 floats)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)

    def w(self, key):
        """The copy of a kernel the GEMMs read (tf32-rounded shadow when enabled)."""
        return self.view(self.shadow if self.shadow is not None else self.params, key)

    def view(self, buf, key):
        off, shape = self.layout[key]
        return buf[off:off + int(np.prod(shape))].view(*shape)

    def p(self, key): return self.view(self.params, key)
    def g(self, key): return self.view(self.grads, key)

    def g_kernel_and_bias(self, layer):
        """Gradient of ``layer``'s kernel AND bias as one (in + 1, out) matrix: the bias follows its kernel in
        the flat buffer, so the last row of  [x | 1]^T dY  lands exactly on the bias gradient."""
        off, (fi, fo) = self.layout[layer + "/kernel"]
        boff, _ = self.layout[layer + "/bias"]
        if boff != off + fi * fo:
            raise ValueError(f"{layer}: bias does not follow the kernel contiguously (in*out % 4 != 0)")
        return self.grads[off:off + (fi + 1) * fo].view(fi + 1, fo)

    def num_parameters(self):
        return sum(int(np.prod(s)) for _, s in self.layout.values())

    def load_dict(self, params: dict):
        for k, (off, shape) in self.layout.items():
            t = torch.as_tensor(np.asarray(params[k]), dtype=torch.float32).reshape(shape)
            self.p(k).copy_(t)
        self.sync_shadow()

    def to_dict(self, buf=None) -> dict:
        buf = self.params if buf is None else buf
        return {k: self.view(buf, k).detach().cpu().numpy().copy() for k in self.layout}

    def init_glorot(self, seed=0):
        """Keras defaults: glorot_uniform kernels, zero biases (SURVEY.md §8d seeding)."""
        from ..synth import glorot_uniform
        gen = torch.Generator().manual_seed(seed)
        out = {}
        for name, fi, fo in layer_specs(self.num_cards):
            out[name + "/kernel"] = glorot_uniform(fi, fo, gen)
            out[name + "/bias"] = np.zeros(fo, dtype=np.float32)
        self.load_dict(out)


# ------------------------------------------------------------------ kernel wrappers
def gemm(a, b, c, *, transa=False, transb=False, bias=None, relu=False, mask=None, accumulate=False,
         precision="fp32", round_out=False):
    """c = epi(op(a) @ op(b)) on 2-D row-major float32 tensors (strides = leading dims)."""
    m, n = c.shape
    k = a.shape[0] if transa else a.shape[1]
    if precision == "fp32":
        call("cc_gemm_f32_simt", int(transa), int(transb), m, n, k, ptr(a), a.stride(0), ptr(b), b.stride(0),
             ptr(c), c.stride(0), ptr(bias), int(relu), ptr(mask), mask.stride(0) if mask is not None else 0,
             int(accumulate), stream_ptr())
    else:
        from . import tensorcore
        tensorcore.gemm(a, b, c, transa=transa, transb=transb, bias=bias, relu=relu, mask=mask,
                        accumulate=accumulate, precision=precision, round_out=round_out)
    return c


def colsum(x, out, ws, accumulate=False):
    m, n = x.shape
    call("cc_colsum_f32", ptr(x), x.stride(0), m, n, ptr(ws), ptr(out), int(accumulate), stream_ptr())


def bag_fwd(w, idx, row_start, row_len, bias, out, relu=True, round_tf32=False):
    call("cc_bag_fwd", ptr(w), w.stride(0), w.shape[1], ptr(idx), ptr(row_start), ptr(row_len), out.shape[0],
         ptr(bias), ptr(out), out.stride(0), int(relu), int(round_tf32), stream_ptr())


def bag_bwd(g, idx, row_start, row_len, dw):
    call("cc_bag_bwd", ptr(g), g.stride(0), dw.shape[1], ptr(idx), ptr(row_start), ptr(row_len), g.shape[0],
         ptr(dw), dw.stride(0), stream_ptr())


class SparseBatch:
    """Index lists on the device: row b = idx[row_start[b] : row_start[b] + row_len[b]]."""

    def __init__(self, idx, row_start, row_len):
        self.idx, self.row_start, self.row_len = idx, row_start, row_len

    @property
    def batch(self): return self.row_len.shape[0]

    @classmethod
    def from_csr(cls, csr: CubeCSR, device):
        idx = torch.from_numpy(np.ascontiguousarray(csr.indices, dtype=np.int32)).to(device)
        start = torch.from_numpy(np.ascontiguousarray(csr.indptr[:-1], dtype=np.int64)).to(device)
        ln = torch.from_numpy(np.diff(csr.indptr).astype(np.int32)).to(device)
        if idx.numel() == 0:
            idx = torch.zeros(1, dtype=torch.int32, device=device)
        return cls(idx, start, ln)

    @classmethod
    def from_rows(cls, rows: torch.Tensor):
        n = rows.shape[0]
        return cls(rows, torch.arange(n, dtype=torch.int64, device=rows.device),
                   torch.ones(n, dtype=torch.int32, device=rows.device))


def _as_sparse(x, num_cards, device) -> SparseBatch:
    if isinstance(x, SparseBatch):
        return x
    if isinstance(x, CubeCSR):
        return SparseBatch.from_csr(x, device)
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    x = np.asarray(x)
    if x.ndim == 1:
        x = x[None, :]
    if x.shape[1] != num_cards:
        raise ValueError(f"expected (batch, {num_cards}) cube vectors, got {x.shape}")
    if not np.isin(x, (0, 1)).all():
        raise ValueError("the sparse first layer needs binary cube vectors")
    return SparseBatch.from_csr(CubeCSR.from_dense(x), device)


class _Tower:
    """``model.encoder`` / ``model.decoder`` / ``model.decoder_for_reg`` callables."""

    def __init__(self, model, kind):
        self.model, self.kind = model, kind

    def __call__(self, x, training=None):
        m = self.model
        if self.kind == "encoder":
            return m._encode(_as_sparse(x, m.N, m.device))
        h = torch.as_tensor(x, dtype=torch.float32, device=m.device) if not isinstance(x, torch.Tensor) else x.to(m.device)
        z = m._decode(h, "main" if self.kind == "decoder" else "reg")
        if self.kind == "decoder":
            out = torch.empty_like(z)
            call("cc_sigmoid_f32", ptr(z), ptr(out), z.numel(), stream_ptr())
            return out
        return torch.softmax(z, dim=1)   # API-parity path only; the train step uses the fused loss kernel


class CC_Recommender:
    """Reference ``model.py:82-125``: encoder E, sigmoid decoder D1, softmax decoder D2."""

    def __init__(self, num_cards, device="cuda", seed=0, precision="fp32"):
        _lib.load()
        self.N = int(num_cards)
        self.device = torch.device(device)
        self.precision = precision
        self.store = ParamStore(self.N, self.device)
        # "fp32": exact FFMA GEMMs; "tf32": tcgen05 kind::tf32 everywhere; "bf16": the seven 512 <-> C passes of the
        # train step on kind::f16 with bf16 operands (fp32 accumulation, fp32 master weights), everything else as tf32
        if precision in ("tf32", "bf16"):
            self.store.enable_tf32_shadow()
        elif precision not in ("fp32",):
            raise NotImplementedError(f"precision {precision!r}: supported modes are 'fp32', 'tf32' and 'bf16'")
        self.gemm_precision = "tf32" if precision == "bf16" else precision      # model-level (inference) GEMMs
        self.store.init_glorot(seed)
        self.encoder = _Tower(self, "encoder")
        self.decoder = _Tower(self, "decoder")
        self.decoder_for_reg = _Tower(self, "reg")

    # -- pieces ---------------------------------------------------------------
    def _encode(self, sb: SparseBatch) -> torch.Tensor:
        s = self.store
        b = sb.batch
        h = torch.empty((b, 512), dtype=torch.float32, device=self.device)
        rnd = self.precision != "fp32"
        bag_fwd(s.p("encoder_e1/kernel"), sb.idx, sb.row_start, sb.row_len, s.p("encoder_e1/bias"), h, round_tf32=rnd)
        if self._small_chain():       # the three layers in one launch (cc_chain_tc): same bits, two launches fewer
            from . import tensorcore
            outs = [torch.empty((b, width), dtype=torch.float32, device=self.device) for width in HIDDEN[1:]]
            tensorcore.chain(h, [(s.w(name + "/kernel"), True, s.p(name + "/bias"), None, o)
                                 for name, o in zip(ENC_NAMES[1:], outs)], relu=True, round_out=rnd)
            return outs[-1]
        for name, width in zip(ENC_NAMES[1:], HIDDEN[1:]):
            o = torch.empty((b, width), dtype=torch.float32, device=self.device)
            gemm(h, s.w(name + "/kernel"), o, bias=s.p(name + "/bias"), relu=True, precision=self.gemm_precision,
                 round_out=rnd)
            h = o
        return h

    def _small_chain(self):
        import os
        return self.precision != "fp32" and self.gemm_precision == "tf32" and os.environ.get("CC_SMALL_CHAIN", "1") != "0"

    def _decode(self, h: torch.Tensor, prefix: str) -> torch.Tensor:
        """Logits (batch, C) of decoder ``prefix`` ('main' | 'reg')."""
        s = self.store
        b = h.shape[0]
        names = dec_names(prefix)
        rnd = self.precision != "fp32"
        h = h.contiguous()
        if rnd:   # a caller-supplied latent is an operand of a kind::tf32 GEMM: round it like every other one
            hr = torch.empty_like(h)
            call("cc_round_tf32", ptr(h), ptr(hr), h.numel(), stream_ptr())
            h = hr
        if self._small_chain():
            from . import tensorcore
            outs = [torch.empty((b, width), dtype=torch.float32, device=self.device) for width in (128, 256, 512)]
            tensorcore.chain(h, [(s.w(name + "/kernel"), True, s.p(name + "/bias"), None, o)
                                 for name, o in zip(names[:3], outs)], relu=True, round_out=rnd)
            h = outs[-1]
        else:
            for name, width in zip(names[:3], (128, 256, 512)):
                o = torch.empty((b, width), dtype=torch.float32, device=self.device)
                gemm(h, s.w(name + "/kernel"), o, bias=s.p(name + "/bias"), relu=True, precision=self.gemm_precision,
                     round_out=rnd)
                h = o
        cpad = (self.N + 3) // 4 * 4
        z = torch.empty((b, cpad), dtype=torch.float32, device=self.device)[:, :self.N]
        gemm(h, s.w(names[3] + "/kernel"), z, bias=s.p(names[3] + "/bias"), precision=self.gemm_precision)
        return z

    def call(self, inputs, training=None):
        """``(x, identity) -> (reconstruction, decoded_for_reg)`` (reference model.py:100-125).
        ``identity`` may be one-hot rows of I or an integer vector of row ids."""
        x, identity = inputs
        rec = self.decoder(self.encoder(x))
        if isinstance(identity, (np.ndarray, torch.Tensor)) and np.asarray(identity.cpu() if isinstance(identity, torch.Tensor) else identity).ndim == 2:
            ident = np.asarray(identity.cpu() if isinstance(identity, torch.Tensor) else identity)
            if not ((ident.sum(1) == 1).all() and np.isin(ident, (0, 1)).all()):
                raise ValueError("identity rows must be one-hot")
            rows = ident.argmax(1)
        else:
            rows = np.asarray(identity.cpu() if isinstance(identity, torch.Tensor) else identity)
        rows_t = torch.as_tensor(rows, dtype=torch.int32, device=self.device)
        reg = self.decoder_for_reg(self._encode(SparseBatch.from_rows(rows_t)))
        return rec, reg

    __call__ = call

    # -- weights ----------------------------------------------------------------
    def get_weights_dict(self):
        return self.store.to_dict()

    def set_weights_dict(self, params):
        self.store.load_dict(params)

    completed_epochs = 0        # epochs of training behind these weights (written by save(), restored by load())

    def save(self, path, epoch=None):
        """Own checkpoint format (npz of the 24 Keras-named tensors + Adam slots + step + completed epochs);
        replaces ``autoencoder.save(dest, save_format='tf')`` (reference train.py:112-115).  The file is written
        under a temporary name and renamed, so an interrupted save never leaves a truncated checkpoint behind."""
        import os
        os.makedirs(path, exist_ok=True)
        if epoch is not None:
            self.completed_epochs = int(epoch)
        blob = {k: v for k, v in self.store.to_dict().items()}
        blob.update({"adam_m/" + k: v for k, v in self.store.to_dict(self.store.adam_m).items()})
        blob.update({"adam_v/" + k: v for k, v in self.store.to_dict(self.store.adam_v).items()})
        blob["step"] = self.store.step.cpu().numpy()
        blob["num_cards"] = np.int64(self.N)
        blob["completed_epochs"] = np.int64(self.completed_epochs)
        tmp = os.path.join(path, "cc_recommender.tmp.npz")
        np.savez(tmp, **blob)
        os.replace(tmp, os.path.join(path, "cc_recommender.npz"))

    @classmethod
    def load(cls, path, device="cuda", precision="fp32"):
        import os
        if not os.path.exists(os.path.join(path, "cc_recommender.npz")):
            if os.path.exists(os.path.join(path, "saved_model.pb")):
                raise FileNotFoundError(
                    f"{path} holds a TensorFlow SavedModel (the reference's format, train.py:112-115), which this package "
                    f"cannot read without TensorFlow.  Convert it once on a machine that has TensorFlow with "
                    f"`python -m cubecobrarecommender_b200.scripts.convert_savedmodel {path} <out_dir>` and load <out_dir>.")
            raise FileNotFoundError(f"no cc_recommender.npz under {path} (save one with CC_Recommender.save)")
        blob = np.load(os.path.join(path, "cc_recommender.npz"))
        model = cls(int(blob["num_cards"]), device=device, precision=precision)
        model.store.load_dict({k: blob[k] for k in model.store.layout})
        for k in model.store.layout:
            if "adam_m/" + k in blob:
                model.store.view(model.store.adam_m, k).copy_(torch.as_tensor(blob["adam_m/" + k]))
                model.store.view(model.store.adam_v, k).copy_(torch.as_tensor(blob["adam_v/" + k]))
        if "step" in blob:
            model.store.step.copy_(torch.as_tensor(blob["step"]))
        if "completed_epochs" in blob:
            model.completed_epochs = int(blob["completed_epochs"])
        return model


def load_model(path, device="cuda", precision="fp32"):
    """Stand-in for ``tensorflow.keras.models.load_model`` (reference ml_recommend.py:54)."""
    return CC_Recommender.load(path, device=device, precision=precision)
