"""Host-side API of the B200-native arch1 step (Python stand-in for the Lua host).

``Arch1Model`` holds what 002_train_vqa_arch1/002_train_baseline.lua keeps in its globals
(embedding_net_q / encoder_net_q / multimodal_net parameters, RMSprop state) and exposes the
step pieces JdJ is made of.  Everything here is plumbing: all arithmetic happens in libnvqa.so.
"""
import ctypes as C
from dataclasses import dataclass, asdict

import numpy as np

from . import _lib
from ._lib import FUSION_AXB, FUSION_ASKIPB  # noqa: F401
from ._lib import (PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16, PREC_BF16X2, BLOCK_ENCODER, BLOCK_EMBEDDING,  # noqa: F401
                   BLOCK_MULTIMODAL, MODE_EVAL, MODE_TRAIN, PHASE_HEAD, PHASE_LSTM, PHASE_EMBED, PHASE_ALL,
                   NvqaError)

__all__ = ["Arch1Config", "Arch1Model", "Arch2Config", "Arch2Model", "synth_params2", "synth_batch2", "BLOCK_CNN",
           "AEConfig", "AEModel", "synth_params_ae", "synth_batch_ae", "BLOCK_AE_ENCODER", "BLOCK_AE_DECODER", "BLOCK_AE_LOOKUP", "DeviceBuffer", "right_align", "pack_batch", "mc_select", "FUSION_AXB", "FUSION_ASKIPB", "synth_batch", "synth_params",
           "device_count", "launch_count", "PREC_FP32_SIMT", "PREC_BF16X3", "PREC_BF16", "PREC_BF16X2",
           "BLOCK_ENCODER", "BLOCK_EMBEDDING", "BLOCK_MULTIMODAL", "MODE_EVAL", "MODE_TRAIN", "PHASE_HEAD",
           "PHASE_LSTM", "PHASE_EMBED", "PHASE_ALL", "NvqaError", "DECAY_FACTOR"]

DECAY_FACTOR = 0.99997592083      # 002_train_baseline.lua:78
BLOCK_CNN = 0                     # arch2 blocks: cnn_w, encoder_w_q (LSTM core + LookupTable), multimodal_w
BLOCK_AE_ENCODER, BLOCK_AE_DECODER, BLOCK_AE_LOOKUP = 0, 1, 2   # nn.AutoEncoder:parameters() order


@dataclass
class Arch1Config:
    """Defaults = the cmd:option defaults of 002_train_baseline.lua:22-48 + vocab_oracle.json size."""
    V: int = 14773
    E: int = 200
    H: int = 512
    L: int = 2
    I: int = 4096
    C: int = 1024
    O: int = 1000
    T: int = 26
    B: int = 500
    dropout: float = 0.5
    img_norm: int = 1

    @property
    def S(self):
        return 2 * self.L * self.H


@dataclass
class Arch2Config:
    """Defaults = the cmd:option defaults of 003_train_vqa_arch2/002_train_baseline.lua:30-43 (I = 2048 for the
    Inception features of BASELINE config 4)."""
    V: int = 14773
    E: int = 512
    H: int = 512
    L: int = 1
    I: int = 4096
    O: int = 1000
    T: int = 26
    B: int = 500
    dropout: float = 0.5
    img_norm: int = 1
    C: int = 0

    @property
    def S(self):
        return self.H


def device_count():
    return _lib.load().nvqa_device_count()


def launch_count():
    return int(_lib.load().nvqa_launch_count())


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def right_align(seq, lengths):
    """misc/RNNUtils.lua:54-61 via the C ABI."""
    seq, lengths = _i32(seq), _i32(lengths)
    out = np.empty_like(seq)
    _lib.check(_lib.load().nvqa_right_align(seq.ctypes.data_as(_lib.c_i32p), lengths.ctypes.data_as(_lib.c_i32p),
                                            seq.shape[0], seq.shape[1], out.ctypes.data_as(_lib.c_i32p)))
    return out


def pack_batch(q_ra, lengths):
    """sort_encoding_onehot_right_align (misc/RNNUtils.lua:84-125) minus the one-hot: returns
    (words, batch_sizes, sort_index, sort_index_inverse), permutations 1-based like Torch."""
    q_ra, lengths = _i32(q_ra), _i32(lengths)
    B, T = q_ra.shape
    words = np.zeros(B * T, dtype=np.int32)
    sizes = np.zeros(T, dtype=np.int32)
    sidx = np.zeros(B, dtype=np.int32)
    inv = np.zeros(B, dtype=np.int32)
    nw, ns = C.c_int32(0), C.c_int32(0)
    p = lambda a: a.ctypes.data_as(_lib.c_i32p)
    _lib.check(_lib.load().nvqa_pack_batch(p(q_ra), p(lengths), B, T, p(words), p(sizes), p(sidx), p(inv),
                                           C.byref(nw), C.byref(ns)))
    return words[:nw.value].copy(), sizes[:ns.value].copy(), sidx, inv


def mc_select(scores, mc_ids):
    """004_eval_model.lua:257-271 through the C ABI: scores [n x O] float32, mc_ids [n x K] int32 (0 = padding)."""
    scores, mc_ids = _f32(scores), _i32(mc_ids)
    out = np.empty(scores.shape[0], dtype=np.int32)
    _lib.check(_lib.load().nvqa_mc_select(scores.ctypes.data_as(_lib.c_f32p), mc_ids.ctypes.data_as(_lib.c_i32p),
                                          scores.shape[0], scores.shape[1], mc_ids.shape[1], out.ctypes.data_as(_lib.c_i32p)))
    return out


class DeviceBuffer:
    """A cudaMalloc'ed array owned by Python (plumbing for tests / bench)."""

    def __init__(self, model, array):
        self.model = model
        self.array = np.ascontiguousarray(array)
        self.ptr = C.c_void_p()
        _lib.check(_lib.load().nvqa_device_alloc(C.byref(self.ptr), self.array.nbytes))
        _lib.check(_lib.load().nvqa_memcpy_h2d(model.handle, self.ptr, _ptr(self.array), self.array.nbytes))

    def get(self):
        out = np.empty_like(self.array)
        _lib.check(_lib.load().nvqa_memcpy_d2h(self.model.handle, _ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            _lib.load().nvqa_device_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Arch1Model:
    ARCH = 1
    CONFIG = Arch1Config

    def __init__(self, cfg=None, precision=PREC_FP32_SIMT, device=0, **overrides):
        self.lib = _lib.load()
        cfg = cfg or self.CONFIG()
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.precision = precision
        c = _lib.nvqa_config(arch=self.ARCH, precision=precision, device=device, **{k: v for k, v in asdict(cfg).items()})
        self.handle = C.c_void_p()
        _lib.check(self.lib.nvqa_model_create(C.byref(c), C.byref(self.handle)))
        self._keep = []

    def close(self):
        if getattr(self, "handle", None):
            self.lib.nvqa_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters (Torch flat layout on the host side) ----
    def param_count(self, block):
        n = C.c_int64(0)
        _lib.check(self.lib.nvqa_param_count(self.handle, block, C.byref(n)))
        return n.value

    def _get(self, fn, block):
        out = np.empty(self.param_count(block), dtype=np.float32)
        _lib.check(fn(self.handle, block, out.ctypes.data_as(_lib.c_f32p)))
        return out

    def set_params(self, block, w):
        w = _f32(w)
        assert w.size == self.param_count(block)
        _lib.check(self.lib.nvqa_params_set(self.handle, block, w.ctypes.data_as(_lib.c_f32p)))

    def get_params(self, block):
        return self._get(self.lib.nvqa_params_get, block)

    def get_grads(self, block):
        return self._get(self.lib.nvqa_grads_get, block)

    def get_rms(self, block):
        return self._get(self.lib.nvqa_rms_get, block)

    def device_views(self):
        p, g = C.c_void_p(), C.c_void_p()
        off = (C.c_int64 * 4)()
        _lib.check(self.lib.nvqa_device_views(self.handle, C.byref(p), C.byref(g), off))
        return p.value, g.value, list(off)

    # ---- batch ----
    def set_batch_host(self, q_ra, lengths, fc7, labels=None):
        q_ra, lengths, fc7 = _i32(q_ra), _i32(lengths), _f32(fc7)
        labels = None if labels is None else _i32(labels)
        self._keep = [q_ra, lengths, fc7, labels]
        _lib.check(self.lib.nvqa_set_batch_host(self.handle, _ptr(q_ra), _ptr(lengths), _ptr(fc7),
                                                None if labels is None else _ptr(labels), q_ra.shape[0]))
        self.sync()

    def set_batch_device(self, q, lengths, fc7, labels, B):
        """q, lengths, fc7, labels: DeviceBuffer (labels may be None)."""
        self._keep = [q, lengths, fc7, labels]
        _lib.check(self.lib.nvqa_set_batch(self.handle, q.ptr, lengths.ptr, fc7.ptr,
                                           None if labels is None else labels.ptr, B))

    def set_masks(self, emb=None, lstm=None, q=None, i=None, z=None):
        """Explicit Dropout multipliers in the padded layout (see include/nvqa.h); None clears."""
        bufs = [None if a is None else DeviceBuffer(self, _f32(a)) for a in (emb, lstm, q, i, z)]
        self._masks = bufs
        _lib.check(self.lib.nvqa_set_masks(self.handle, *[None if b is None else b.ptr for b in bufs]))

    # ---- step pieces ----
    def forward(self, mode=MODE_EVAL, seed=0):
        _lib.check(self.lib.nvqa_forward(self.handle, mode, seed))

    def loss(self):
        out = C.c_float(0)
        _lib.check(self.lib.nvqa_loss(self.handle, C.byref(out)))
        return out.value

    def backward(self, phase=PHASE_ALL):
        _lib.check(self.lib.nvqa_backward(self.handle, phase))

    def set_variant(self, fusion=0, lr_scale=1.0, norm_split=0):
        """003_train_ae_based_wp.lua (AskipB, -lr_scale) / 003_train_ae_based_ef.lua (two-block image norm)."""
        _lib.check(self.lib.nvqa_set_variant(self.handle, fusion, lr_scale, norm_split))

    def rmsprop_step(self, lr, alpha=0.99, eps=1e-8, wd=0.0, clamp=10.0, grad_scale=1.0):
        _lib.check(self.lib.nvqa_rmsprop_step(self.handle, lr, alpha, eps, wd, clamp, grad_scale))

    def scores(self, B):
        out = np.empty((B, self.cfg.O), dtype=np.float32)
        _lib.check(self.lib.nvqa_scores_get(self.handle, out.ctypes.data_as(_lib.c_f32p)))
        return out

    def argmax(self, B):
        out = np.empty(B, dtype=np.int32)
        _lib.check(self.lib.nvqa_argmax_get(self.handle, out.ctypes.data_as(_lib.c_i32p)))
        return out

    def state(self, B):
        out = np.empty((B, self.cfg.S), dtype=np.float32)
        _lib.check(self.lib.nvqa_state_get(self.handle, out.ctypes.data_as(_lib.c_f32p)))
        return out

    def sync(self):
        _lib.check(self.lib.nvqa_sync(self.handle))

    def train_step(self, lr, seed):
        """JdJ + clamp + optimizer on the batch already set (device-resident twin of train_step_host)."""
        _lib.check(self.lib.nvqa_train_step(self.handle, lr, seed))

    # ---- fused convenience (host buffers in, scalar out) ----
    def train_step_host(self, q_ra, lengths, fc7, labels, lr, seed):
        out = C.c_float(0)
        q_ra, lengths, fc7, labels = _i32(q_ra), _i32(lengths), _f32(fc7), _i32(labels)   # no-ops for conforming arrays
        _lib.check(self.lib.nvqa_train_step_host(self.handle, _ptr(q_ra), _ptr(lengths), _ptr(fc7), _ptr(labels),
                                                 q_ra.shape[0], lr, seed, C.byref(out)))
        return out.value

    def eval_step_host(self, q_ra, lengths, fc7):
        q_ra, lengths, fc7 = _i32(q_ra), _i32(lengths), _f32(fc7)
        ans = np.empty(q_ra.shape[0], dtype=np.int32)
        _lib.check(self.lib.nvqa_eval_step_host(self.handle, _ptr(q_ra), _ptr(lengths), _ptr(fc7), q_ra.shape[0],
                                                _ptr(ans)))
        return ans


class Arch2Model(Arch1Model):
    """003_train_vqa_arch2: cnn_projection + nn.Encoder (image, START, words through a LookupTable LSTM) + head.
    Blocks: BLOCK_CNN (0), BLOCK_EMBEDDING slot = encoder_w_q (1), BLOCK_MULTIMODAL (2).  Questions are passed as
    stored (left-aligned, zero-padded)."""
    ARCH = 2
    CONFIG = Arch2Config

    def set_steps(self, steps):
        _lib.check(self.lib.nvqa_set_steps(self.handle, steps))

    def set_lookup_grad_literal(self, on=True):
        _lib.check(self.lib.nvqa_set_lookup_grad_literal(self.handle, 1 if on else 0))

    def set_stale_h0_literal(self, on=True):
        """the literal reference's stale-gradient initial state (misc/Encoder_lstm.lua:238-239)."""
        _lib.check(self.lib.nvqa_set_stale_h0_literal(self.handle, 1 if on else 0))

    def set_masks(self, lstm=None, z=None):
        super().set_masks(emb=None, lstm=lstm, q=None, i=None, z=z)


@dataclass
class AEConfig:
    """Defaults = cmd:option defaults of 001_train_autoencoder/001_train_arch1_text_autoencoder.lua:27-37 and the
    prepro defaults (000_prepro_book_corpus.py:265,270): V = 20000 (+1), seq_length 16, batch 1000."""
    V: int = 20000
    E: int = 512
    H: int = 512
    L: int = 1
    T: int = 16
    B: int = 1000
    dropout: float = 0.5
    I: int = 4
    C: int = 0
    O: int = 4
    img_norm: int = 0

    @property
    def S(self):
        return 2 * self.L * self.H


class AEModel(Arch1Model):
    """001_train_autoencoder: nn.AutoEncoder (LSTM encoder -> LSTM decoder over a shared LookupTable) +
    nn.LanguageModelCriterion + clamp / weight decay / adam.  Blocks: BLOCK_AE_ENCODER, BLOCK_AE_DECODER (LSTM core then
    Linear(H, V+1)), BLOCK_AE_LOOKUP.  seq is [B x T], zero-padded on the right (the reference's [T x B] transposed)."""
    ARCH = 3
    CONFIG = AEConfig

    def set_batch_host(self, seq, lengths=None):
        seq = _i32(seq)
        lengths = _i32((seq != 0).sum(axis=1) if lengths is None else lengths)
        self._keep = [seq, lengths]
        _lib.check(self.lib.nvqa_set_batch_host(self.handle, _ptr(seq), _ptr(lengths), None, None, seq.shape[0]))
        self.sync()

    def set_batch_device(self, seq, B, tmax):
        self._keep = [seq]
        _lib.check(self.lib.nvqa_set_batch(self.handle, seq.ptr, None, None, None, B))
        _lib.check(self.lib.nvqa_set_steps(self.handle, tmax))

    def set_lookup_grad_literal(self, on=True):
        _lib.check(self.lib.nvqa_set_lookup_grad_literal(self.handle, 1 if on else 0))

    def adam_step(self, lr=1e-5, beta1=0.8, beta2=0.999, eps=1e-8, wd=1e-6, clamp=0.1, grad_scale=1.0):
        _lib.check(self.lib.nvqa_adam_step(self.handle, lr, beta1, beta2, eps, wd, clamp, grad_scale))

    def logprobs(self, step, B):
        out = np.empty((B, self.cfg.V + 1), dtype=np.float32)
        _lib.check(self.lib.nvqa_logprobs_get(self.handle, step, out.ctypes.data_as(_lib.c_f32p)))
        return out

    def train_step_host(self, seq, lengths, lr, seed):
        out = C.c_float(0)
        seq, lengths = _i32(seq), _i32(lengths)
        _lib.check(self.lib.nvqa_train_step_host(self.handle, _ptr(seq), _ptr(lengths), None, None, seq.shape[0], lr, seed,
                                                 C.byref(out)))
        return out.value


def synth_params_ae(cfg, seed=123):
    """uniform(-0.08, 0.08) over encoder, decoder, lookup_table (flat parameter order of nn.AutoEncoder)."""
    r = np.random.default_rng(seed)
    n_core = sum(4 * cfg.H * ((cfg.E if l == 0 else cfg.H) + cfg.H + 2) for l in range(cfg.L))
    enc = r.uniform(-0.08, 0.08, n_core).astype(np.float32)
    dec = r.uniform(-0.08, 0.08, n_core + (cfg.V + 1) * cfg.H + cfg.V + 1).astype(np.float32)
    lut = r.uniform(-0.08, 0.08, (cfg.V + 1) * cfg.E).astype(np.float32)
    return enc, dec, lut


def synth_batch_ae(cfg, B, seed=123, min_len=4):
    """book-corpus-shaped token sequences: U{1..V} tokens, lengths U{min_len..T}, zero-padded right."""
    r = np.random.default_rng(seed)
    T = cfg.T
    lengths = r.integers(min_len, T + 1, B).astype(np.int32)
    tok = r.integers(1, cfg.V + 1, (B, T)).astype(np.int32)
    seq = np.where(np.arange(T)[None, :] < lengths[:, None], tok, 0).astype(np.int32)
    return seq, lengths


def synth_params2(cfg, seed=123):
    """uniform(-0.08, 0.08) over cnn_w, encoder_w_q, multimodal_w (003_train_vqa_arch2/002_train_baseline.lua:182-189)."""
    r = np.random.default_rng(seed)
    n_enc = sum(4 * cfg.H * ((cfg.E if l == 0 else cfg.H) + cfg.H + 2) for l in range(cfg.L)) + (cfg.V + 1) * cfg.E
    cnn = r.uniform(-0.08, 0.08, cfg.E * cfg.I + cfg.E).astype(np.float32)
    enc = r.uniform(-0.08, 0.08, n_enc).astype(np.float32)
    mm = r.uniform(-0.08, 0.08, cfg.O * cfg.H + cfg.O).astype(np.float32)
    return cnn, enc, mm


def synth_batch2(cfg, B, seed=123, min_len=None):
    """arch2 batch: questions left-aligned and zero-padded as in data_prepro.h5 (no right_align)."""
    r = np.random.default_rng(seed)
    T = cfg.T
    lengths = np.full(B, T, dtype=np.int32) if min_len is None else r.integers(min_len, T + 1, B).astype(np.int32)
    tok = r.integers(1, cfg.V + 1, (B, T)).astype(np.int32)
    q = np.where(np.arange(T)[None, :] < lengths[:, None], tok, 0).astype(np.int32)
    fc7 = np.maximum(0.0, r.standard_normal((B, cfg.I))).astype(np.float32)
    labels = r.integers(1, cfg.O + 1, B).astype(np.int32)
    return q, lengths, fc7, labels


# ---- synthetic data of the BASELINE.json shape ----------------------------------------
def synth_params(cfg, seed=123):
    """uniform(-0.08, 0.08) over the flat blocks in the order embedding, encoder, multimodal
    (002_train_baseline.lua:174-181).  Returns (enc_w, emb_w, mm_w) in Torch flat layout."""
    r = np.random.default_rng(seed)
    n_enc = sum(4 * cfg.H * ((cfg.E if l == 0 else cfg.H) + cfg.H + 2) for l in range(cfg.L))
    n_emb = cfg.V * cfg.E + cfg.E
    n_mm = cfg.C * cfg.S + cfg.C + cfg.C * cfg.I + cfg.C + cfg.O * cfg.C + cfg.O
    emb = r.uniform(-0.08, 0.08, n_emb).astype(np.float32)
    enc = r.uniform(-0.08, 0.08, n_enc).astype(np.float32)
    mm = r.uniform(-0.08, 0.08, n_mm).astype(np.float32)
    return enc, emb, mm


def synth_batch(cfg, B, seed=123, min_len=None):
    """Random question tokens U{1..V}, lengths = T (or U{min_len..T}), fc7 = max(0, N(0,1)) (post-ReLU
    like VGG fc7), labels U{1..O}.  Returns (q_right_aligned, lengths, fc7_raw, labels)."""
    r = np.random.default_rng(seed)
    T = cfg.T
    lengths = np.full(B, T, dtype=np.int32) if min_len is None else r.integers(min_len, T + 1, B).astype(np.int32)
    q = np.zeros((B, T), dtype=np.int32)
    tok = r.integers(1, cfg.V + 1, (B, T)).astype(np.int32)
    for b in range(B):
        q[b, T - lengths[b]:] = tok[b, :lengths[b]]
    fc7 = np.maximum(0.0, r.standard_normal((B, cfg.I))).astype(np.float32)
    labels = r.integers(1, cfg.O + 1, B).astype(np.int32)
    return q, lengths, fc7, labels
