"""Host-side mirror of the reference's Torch7 module / function interface for the arch1 hot path.

The reference is Lua; no Lua runtime exists here, so this Python module plays the role of the Lua shims
under ``lua/``: SAME names, argument meaning and error behaviour as the reference's own call sites, every call landing
in libnvqa.so through the C ABI.  It lets the parity tests read like the reference's scripts:

    encoder_net_q = LSTM.lstm_conventional(E, H, 1, n, 0.5)             # misc/LSTM.lua:12
    states = rnn_forward(buffer, init_state, inputs, sizes)             # misc/RNNUtils.lua:128
    scores = multimodal_net:forward({tv_q, fv_im})                      # misc/netdef.lua:6 + Linear
    f = criterion:forward(scores, labels)                               # nn.CrossEntropyCriterion
    optim.rmsprop(JdJ, x, config, state)                                # 002_train_baseline.lua:408

Tensors are NumPy arrays on the host (the mirror copies them to the device for each call -- it is an interface shim,
the throughput path is Arch1Model.train_step_host / dp.train_step).  Errors surface as NvqaError, the analogue of the
Lua ``error()`` raised by TH/THNN.

Module protocol (what 002_train_baseline.lua:272-335 calls): ``forward``, ``backward``, ``getParameters`` -> (w, dw),
``zeroGradParameters``, ``training`` / ``evaluate``, ``clone``; plus ``dupe_rnn`` / ``rnn_forward`` / ``rnn_backward``
(misc/RNNUtils.lua:66-81,128-211).  ``build_arch1_nets`` creates the four modules of :139-157 over ONE libnvqa model,
which is what lua/misc/LSTM.lua + lua/misc/netdef.lua do on the Lua side (tests/test_mirror_gpu.py runs the
reference-shaped JdJ through them).
"""
import ctypes as C

import numpy as np

from . import _lib
from .api import (Arch1Config, Arch1Model, DeviceBuffer, BLOCK_ENCODER, BLOCK_EMBEDDING, BLOCK_MULTIMODAL, pack_batch,
                  PREC_FP32_SIMT)
from .api import right_align as _right_align


# ---- misc/RNNUtils.lua ------------------------------------------------------------------------------------------
def join_vector(tensor_table):
    """RNNUtils.lua:22-24."""
    return np.concatenate([np.asarray(t).ravel() for t in tensor_table]).copy()


def split_vector(w, sizes):
    """RNNUtils.lua:25-39: views of w by sizes."""
    out, off = [], 0
    for n in sizes:
        out.append(w[off:off + int(n)])
        off += int(n)
    return out


def right_align(seq, lengths):
    """RNNUtils.lua:54-61."""
    return _right_align(seq, lengths)


def inverse_mapping(ind):
    """RNNUtils.lua:13-16 (1-based in, 1-based out)."""
    return np.argsort(np.asarray(ind), kind="stable") + 1


def sort_encoding_onehot_right_align(batch_word_right_align, batch_length, vocabulary_size):
    """RNNUtils.lua:84-125.  Element [0] is the packed word-id vector, not a dense one-hot [N x V] tensor (the one-hot
    nn.Linear runs as a gather inside the library); [1] batch_sizes, [2] sort_index, [3] sort_index_inverse (1-based)."""
    words, sizes, sidx, inv = pack_batch(batch_word_right_align, batch_length)
    if words.size and (words.min() < 1 or words.max() > vocabulary_size):
        raise _lib.NvqaError("sort_encoding_onehot_right_align: word id out of range")
    return [words.astype(np.int64), sizes.astype(np.int64), sidx.astype(np.int64), inv.astype(np.int64)]


# ---- misc/LSTM.lua ------------------------------------------------------------------------------------------------
class _Module:
    """What every mirrored nn module shares: a flat (w, dw) pair on the host and ONE block of a libnvqa model."""
    BLOCK = None

    def _init_params(self, model, w=None):
        self._model = model
        self._w = np.zeros(model.param_count(self.BLOCK), dtype=np.float32) if w is None else w
        self._dw = np.zeros_like(self._w)
        self.train = True

    def getParameters(self):
        """flat (w, dw): written by the caller (w) / by backward (dw), like Torch's flattened views"""
        return self._w, self._dw

    def parameters(self):
        return [self._w], [self._dw]

    def training(self):
        self.train = True

    def evaluate(self):
        self.train = False

    def zeroGradParameters(self):
        self._dw[...] = 0

    def _push(self):
        self._model.set_params(self.BLOCK, self._w)

    def _backward_block(self, call):
        """runs one *_backward entry point on a zeroed device gradient block and accumulates the result into dw
        (accGradParameters: dw += this call's contribution)"""
        m = self._model
        _lib.check(m.lib.nvqa_grads_zero(m.handle, self.BLOCK))
        call()
        self._dw += m.get_grads(self.BLOCK)


class _LSTMCell(_Module):
    """nngraph gModule returned by LSTM.lstm_conventional: forward({state, x}) -> state' on a packed
    [c1 h1 c2 h2 ...] state (misc/LSTM.lua:15-72); backward({state, x}, dstate') -> {dstate, dx}."""
    BLOCK = BLOCK_ENCODER

    def __init__(self, input_size, rnn_size, n, dropout, model=None, w=None):
        self.input_size, self.rnn_size, self.n, self.dropout = input_size, rnn_size, n, dropout
        self._init_params(model or Arch1Model(Arch1Config(E=input_size, H=rnn_size, L=n, V=8, I=4, C=4, O=4, T=8, B=512,
                                                          dropout=dropout)), w)
        self.masks = None          # explicit Dropout multipliers [(n-1) x rows x H] for training-mode parity

    def clone(self):
        """net:clone() as used by dupe_rnn (misc/RNNUtils.lua:66-81): own parameter / gradient storage, same device model"""
        c = _LSTMCell(self.input_size, self.rnn_size, self.n, self.dropout, self._model, self._w.copy())
        c.train = self.train
        return c

    def _check(self, state, x):
        if state.shape[1] != 2 * self.n * self.rnn_size or x.shape[1] != self.input_size or state.shape[0] != x.shape[0]:
            raise _lib.NvqaError("size mismatch")                      # what THNN's Linear would raise

    def _mask_buffer(self):
        if not (self.train and self.dropout > 0 and self.n > 1):
            return None
        if self.masks is None:
            raise _lib.NvqaError("training-mode forward needs explicit .masks (Torch7's RNG is not reproducible)")
        return DeviceBuffer(self._model, np.ascontiguousarray(self.masks, dtype=np.float32))

    def forward(self, inputs):
        state, x = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        self._check(state, x)
        m = self._model
        self._push()
        mk = self._mask_buffer()
        S_, X_, O_ = DeviceBuffer(m, state), DeviceBuffer(m, x), DeviceBuffer(m, np.zeros_like(state))
        _lib.check(m.lib.nvqa_lstm_cell_forward(m.handle, S_.ptr, X_.ptr, None if mk is None else mk.ptr, state.shape[0], O_.ptr))
        self.output = O_.get()
        return self.output

    def backward(self, inputs, gradOutput):
        """clone:backward({state, x}, dstate') -> {dstate, dx}; dw += this clone's parameter gradients"""
        state, x = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        g = np.ascontiguousarray(gradOutput, dtype=np.float32)
        self._check(state, x)
        if g.shape != state.shape:
            raise _lib.NvqaError("size mismatch")
        m = self._model
        self._push()
        mk = self._mask_buffer()
        S_, X_, G_ = DeviceBuffer(m, state), DeviceBuffer(m, x), DeviceBuffer(m, g)
        DS_, DX_ = DeviceBuffer(m, np.zeros_like(state)), DeviceBuffer(m, np.zeros_like(x))
        self._backward_block(lambda: _lib.check(m.lib.nvqa_lstm_cell_backward(
            m.handle, S_.ptr, X_.ptr, None if mk is None else mk.ptr, G_.ptr, state.shape[0], DS_.ptr, DX_.ptr)))
        self.gradInput = [DS_.get(), DX_.get()]
        return self.gradInput


class LSTM:
    @staticmethod
    def lstm_conventional(input_size, rnn_size, noutput, n, dropout=0):
        """misc/LSTM.lua:12 (noutput is unused by the reference too, :66-68)."""
        return _LSTMCell(input_size, rnn_size, n, dropout)


def dupe_rnn(net, times):
    """misc/RNNUtils.lua:66-81: `times` clones of the cell, each with its own flat (w, dw) (the reference deep-copies the
    prototype; JdJ re-copies encoder_w_q into every clone each iteration, 002_train_baseline.lua:275-280)."""
    return [[net.clone() for _ in range(times)]]


def rnn_forward(net_buffer, init_state, inputs, sizes):
    """RNNUtils.lua:128-154 for right-aligned (non-decreasing) sizes; net_buffer[0] = list of per-step cells.
    Returns the T+1 states like the reference (states[i] = the state fed to step i, grown to that step's size)."""
    cells = net_buffer[0]
    states = [np.asarray(init_state[:int(sizes[0])], dtype=np.float32)]
    for i in range(len(sizes)):
        if i > 0 and sizes[i] > sizes[i - 1]:
            pad = np.array(init_state[:int(sizes[i])], dtype=np.float32)
            pad[:int(sizes[i - 1])] = states[i]
            states[i] = pad
        elif i > 0 and sizes[i] < sizes[i - 1]:
            raise _lib.NvqaError("left-aligned (shrinking) batches are not on the arch1 path")
        states.append(cells[i].forward([states[i], inputs[i]]))
    return states


def rnn_backward(net_buffer, dend_state, doutputs, states, inputs, sizes):
    """RNNUtils.lua:181-210, the branch JdJ takes (doutputs is the dummy output, not a table): walks the clones
    backwards, trimming the state gradient to the rows active at the previous step.  Returns (dstate, dinputs)."""
    cells = net_buffer[0]
    N = len(sizes)
    dstate = np.asarray(dend_state[:int(sizes[N - 1])], dtype=np.float32)
    dinputs = [None] * N
    for i in reversed(range(N)):
        dprev, dx = cells[i].backward([states[i], inputs[i]], dstate)
        dinputs[i] = dx
        dstate = dprev if (i == 0 or sizes[i] == sizes[i - 1]) else dprev[:int(sizes[i - 1])]
    return dstate, dinputs


# ---- embedding_net_q (002_train_baseline.lua:141-144) ---------------------------------------------------------------
class _EmbeddingNet(_Module):
    """nn.Sequential{Linear(V, E), Dropout(0.5), Tanh}: forward takes element [0] of sort_encoding_onehot_right_align
    (the packed word ids standing for the one-hot rows); backward(words, dy) accumulates the Linear's gradients."""
    BLOCK = BLOCK_EMBEDDING

    def __init__(self, model):
        self._init_params(model)
        self.masks = None          # explicit Dropout multipliers [n x E]

    def _mask_buffer(self, n):
        if not (self.train and self._model.cfg.dropout > 0):
            return None
        if self.masks is None:
            raise _lib.NvqaError("training-mode forward needs explicit .masks (Torch7's RNG is not reproducible)")
        return DeviceBuffer(self._model, np.ascontiguousarray(self.masks, dtype=np.float32).reshape(n, -1))

    def forward(self, words):
        w = np.ascontiguousarray(words, dtype=np.int32)
        m = self._model
        if w.size and (w.min() < 1 or w.max() > m.cfg.V):
            raise _lib.NvqaError("size mismatch")                      # a one-hot row wider than Linear(V, E)
        self._push()
        mk = self._mask_buffer(w.size)
        W_, Y_ = DeviceBuffer(m, w), DeviceBuffer(m, np.zeros((w.size, m.cfg.E), np.float32))
        _lib.check(m.lib.nvqa_embedding_forward(m.handle, W_.ptr, None if mk is None else mk.ptr, w.size, Y_.ptr))
        self.output = Y_.get()
        return self.output

    def backward(self, words, gradOutput):
        w = np.ascontiguousarray(words, dtype=np.int32)
        g = np.ascontiguousarray(gradOutput, dtype=np.float32)
        m = self._model
        self._push()
        mk = self._mask_buffer(w.size)
        W_, Y_, G_ = DeviceBuffer(m, w), DeviceBuffer(m, self.output), DeviceBuffer(m, g)
        self._backward_block(lambda: _lib.check(m.lib.nvqa_embedding_backward(
            m.handle, W_.ptr, Y_.ptr, G_.ptr, None if mk is None else mk.ptr, w.size)))
        return None                # the [n x V] gradInput is never used by the reference (:320)


# ---- misc/netdef.lua + multimodal head ----------------------------------------------------------------------------
class _AxB(_Module):
    """netdef.AxB(nhA, nhB, nhcommon, dropout): tanh(Wq drop(q)) (.) tanh(Wi drop(i))   (misc/netdef.lua:6-14).
    Stand-alone it owns the multimodal block of a private model (the Linear(C,O) tail of the block is unused)."""
    BLOCK = BLOCK_MULTIMODAL

    def __init__(self, nhA, nhB, nhcommon, dropout, H, L, model=None):
        self._init_params(model or Arch1Model(Arch1Config(E=4, H=H, L=L, V=8, I=nhB, C=nhcommon, O=4, T=2, B=512,
                                                          dropout=dropout)))
        assert 2 * H * L == nhA
        self.train = False
        self.masks = None          # (mask_q, mask_i)

    def _mask_buffers(self):
        if not self.train:
            return None, None
        if self.masks is None:
            raise _lib.NvqaError("training-mode forward needs explicit .masks")
        return tuple(DeviceBuffer(self._model, np.ascontiguousarray(a, dtype=np.float32)) for a in self.masks[:2])

    def forward(self, inputs):
        q, i = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        m = self._model
        self._push()
        n = q.shape[0]
        Q_, I_ = DeviceBuffer(m, q), DeviceBuffer(m, i)
        O_ = DeviceBuffer(m, np.zeros((n, m.cfg.C), dtype=np.float32))
        mq, mi = self._mask_buffers()
        _lib.check(m.lib.nvqa_axb_forward(m.handle, Q_.ptr, I_.ptr, None if mq is None else mq.ptr,
                                          None if mi is None else mi.ptr, n, O_.ptr))
        self.output = O_.get()
        return self.output

    def backward(self, inputs, gradOutput):
        """-> {dq, di}; dw += the gradients of Wq, bq, Wi, bi"""
        q, i = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        g = np.ascontiguousarray(gradOutput, dtype=np.float32)
        m = self._model
        self._push()
        n = q.shape[0]
        Q_, I_, G_ = DeviceBuffer(m, q), DeviceBuffer(m, i), DeviceBuffer(m, g)
        DQ_, DI_ = DeviceBuffer(m, np.zeros_like(q)), DeviceBuffer(m, np.zeros_like(i))
        mq, mi = self._mask_buffers()
        self._backward_block(lambda: _lib.check(m.lib.nvqa_axb_backward(
            m.handle, Q_.ptr, I_.ptr, None if mq is None else mq.ptr, None if mi is None else mi.ptr, G_.ptr, n,
            DQ_.ptr, DI_.ptr)))
        self.gradInput = [DQ_.get(), DI_.get()]
        return self.gradInput


class _MultimodalNet(_AxB):
    """multimodal_net = nn.Sequential{netdef.AxB(2*H*n, I, C, 0.5), Dropout(0.5), Linear(C, O)}
    (002_train_baseline.lua:151-154).  masks = (mask_q, mask_i, mask_z)."""

    def __init__(self, model):
        c = model.cfg
        _AxB.__init__(self, 2 * c.H * c.L, c.I, c.C, c.dropout, c.H, c.L, model)
        self.train = True

    def _mask3(self):
        mq, mi = self._mask_buffers()
        mz = None if not self.train else DeviceBuffer(self._model, np.ascontiguousarray(self.masks[2], dtype=np.float32))
        return [None if b is None else b.ptr for b in (mq, mi, mz)], (mq, mi, mz)

    def forward(self, inputs):
        q, i = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        m = self._model
        self._push()
        n = q.shape[0]
        Q_, I_ = DeviceBuffer(m, q), DeviceBuffer(m, i)
        O_ = DeviceBuffer(m, np.zeros((n, m.cfg.O), dtype=np.float32))
        ptrs, keep = self._mask3()
        _lib.check(m.lib.nvqa_multimodal_forward(m.handle, Q_.ptr, I_.ptr, *ptrs, n, O_.ptr))
        self.output = O_.get()
        return self.output

    def backward(self, inputs, gradOutput):
        q, i = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        g = np.ascontiguousarray(gradOutput, dtype=np.float32)
        m = self._model
        self._push()
        n = q.shape[0]
        Q_, I_, G_ = DeviceBuffer(m, q), DeviceBuffer(m, i), DeviceBuffer(m, g)
        DQ_, DI_ = DeviceBuffer(m, np.zeros_like(q)), DeviceBuffer(m, np.zeros_like(i))
        ptrs, keep = self._mask3()
        self._backward_block(lambda: _lib.check(m.lib.nvqa_multimodal_backward(
            m.handle, Q_.ptr, I_.ptr, *ptrs, G_.ptr, n, DQ_.ptr, DI_.ptr)))
        self.gradInput = [DQ_.get(), DI_.get()]
        return self.gradInput


class netdef:
    @staticmethod
    def AxB(nhA, nhB, nhcommon, dropout=0, rnn_size=512, rnn_layers=2):
        return _AxB(nhA, nhB, nhcommon, dropout, rnn_size, rnn_layers)


# ---- nn.CrossEntropyCriterion -------------------------------------------------------------------------------------
class CrossEntropyCriterion:
    """criterion:forward(scores, labels) -> number; criterion:backward(scores, labels) -> dscores (sizeAverage)."""

    def __init__(self, num_output, model=None):
        self._model = model or Arch1Model(Arch1Config(E=4, H=4, L=1, V=8, I=4, C=4, O=num_output, T=2, B=1024))
        self._d = None

    def forward(self, scores, labels):
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        if labels.min() < 1 or labels.max() > scores.shape[1]:
            raise _lib.NvqaError("target out of range")                # ClassNLLCriterion's assertion (App. C-10)
        m = self._model
        S_, L_, D_ = DeviceBuffer(m, scores), DeviceBuffer(m, labels), DeviceBuffer(m, np.zeros_like(scores))
        out = C.c_float(0)
        _lib.check(m.lib.nvqa_cross_entropy(m.handle, S_.ptr, L_.ptr, scores.shape[0], C.byref(out), D_.ptr))
        self._d = D_.get()
        self.output = out.value
        return self.output

    def backward(self, scores, labels):
        if self._d is None:
            self.forward(scores, labels)
        return self._d


def build_arch1_nets(cfg, precision=PREC_FP32_SIMT):
    """The model construction of 002_train_baseline.lua:139-157 over ONE libnvqa model:
    returns (embedding_net_q, encoder_net_q, multimodal_net, criterion, model)."""
    model = Arch1Model(cfg, precision=precision)
    return (_EmbeddingNet(model), _LSTMCell(cfg.E, cfg.H, cfg.L, cfg.dropout, model), _MultimodalNet(model),
            CrossEntropyCriterion(cfg.O, model), model)


# ---- optim.rmsprop ------------------------------------------------------------------------------------------------
class optim:
    _model = None

    @staticmethod
    def rmsprop(opfunc, x, config=None, state=None):
        """optim.rmsprop(opfunc, x, config, state) -> x, {fx}: same defaults as Torch's optim (learningRate 1e-2,
        alpha 0.99, epsilon 1e-8, weightDecay 0); state['m'] is created on first use.  x is updated in place."""
        config = config if config is not None else {}
        state = state if state is not None else config
        lr = config.get("learningRate", 1e-2)
        alpha = config.get("alpha", 0.99)
        eps = config.get("epsilon", 1e-8)
        wd = config.get("weightDecay", 0.0)
        fx, dfdx = opfunc(x)
        if "m" not in state:
            state["m"] = np.zeros_like(x, dtype=np.float32)
        if optim._model is None:
            optim._model = Arch1Model(Arch1Config(E=4, H=4, L=1, V=8, I=4, C=4, O=4, T=2, B=2))
        m = optim._model
        X_, G_, M_ = DeviceBuffer(m, np.ascontiguousarray(x, dtype=np.float32)), \
            DeviceBuffer(m, np.ascontiguousarray(dfdx, dtype=np.float32)), DeviceBuffer(m, state["m"])
        # the reference clamps inside JdJ (:329); here no extra clamp (1e30) and no gradient scaling
        _lib.check(m.lib.nvqa_rmsprop_vector(m.handle, X_.ptr, G_.ptr, M_.ptr, x.size, lr, alpha, eps, wd, 1e30, 1.0))
        x[...] = X_.get().reshape(x.shape)
        state["m"][...] = M_.get()
        return x, [fx]
