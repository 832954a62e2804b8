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
"""
import ctypes as C

import numpy as np

from . import _lib
from .api import Arch1Config, Arch1Model, DeviceBuffer, BLOCK_ENCODER, BLOCK_MULTIMODAL, pack_batch
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
class _LSTMCell:
    """nngraph gModule returned by LSTM.lstm_conventional: forward({state, x}) -> state' on a packed
    [c1 h1 c2 h2 ...] state (misc/LSTM.lua:15-72)."""

    def __init__(self, input_size, rnn_size, n, dropout, model=None):
        self.input_size, self.rnn_size, self.n, self.dropout = input_size, rnn_size, n, dropout
        self.train = True
        self._model = model or Arch1Model(Arch1Config(E=input_size, H=rnn_size, L=n, V=8, I=4, C=4, O=4, T=8, B=512,
                                                      dropout=dropout))
        self._w = np.zeros(self._model.param_count(BLOCK_ENCODER), dtype=np.float32)
        self._dw = np.zeros_like(self._w)
        self.masks = None          # explicit Dropout multipliers [(n-1) x rows x H] for training-mode parity

    def getParameters(self):
        """flat (w, dw); call .sync() (or forward) after writing into w"""
        return self._w, self._dw

    def training(self):
        self.train = True

    def evaluate(self):
        self.train = False

    def forward(self, inputs):
        state, x = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        if state.shape[1] != 2 * self.n * self.rnn_size or x.shape[1] != self.input_size:
            raise _lib.NvqaError("size mismatch")                      # what THNN's Linear would raise
        m = self._model
        m.set_params(BLOCK_ENCODER, self._w)
        rows = state.shape[0]
        mk = None
        if self.train and self.dropout > 0 and self.n > 1:
            if self.masks is None:
                raise _lib.NvqaError("training-mode forward needs explicit .masks (Torch7's RNG is not reproducible)")
            mk = DeviceBuffer(m, np.ascontiguousarray(self.masks, dtype=np.float32))
        S_, X_, O_ = DeviceBuffer(m, state), DeviceBuffer(m, x), DeviceBuffer(m, np.zeros_like(state))
        _lib.check(m.lib.nvqa_lstm_cell_forward(m.handle, S_.ptr, X_.ptr, None if mk is None else mk.ptr, rows, O_.ptr))
        self.output = O_.get()
        return self.output


class LSTM:
    @staticmethod
    def lstm_conventional(input_size, rnn_size, noutput, n, dropout=0):
        """misc/LSTM.lua:12 (noutput is unused by the reference too, :66-68)."""
        return _LSTMCell(input_size, rnn_size, n, dropout)


def rnn_forward(net_buffer, init_state, inputs, sizes):
    """RNNUtils.lua:128-154 for right-aligned (non-decreasing) sizes; net_buffer[0] = list of per-step cells (clones
    share one parameter vector here).  Returns the T+1 states like the reference."""
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


# ---- misc/netdef.lua + multimodal head ----------------------------------------------------------------------------
class _AxB:
    """netdef.AxB(nhA, nhB, nhcommon, dropout): tanh(Wq drop(q)) (.) tanh(Wi drop(i))   (misc/netdef.lua:6-14)"""

    def __init__(self, nhA, nhB, nhcommon, dropout, H, L):
        self._model = Arch1Model(Arch1Config(E=4, H=H, L=L, V=8, I=nhB, C=nhcommon, O=4, T=2, B=512, dropout=dropout))
        assert 2 * H * L == nhA
        self.train = False
        n = self._model.param_count(BLOCK_MULTIMODAL)
        self._w = np.zeros(n, dtype=np.float32)
        self.masks = None          # (mask_q, mask_i)

    def getParameters(self):
        return self._w, np.zeros_like(self._w)

    def training(self):
        self.train = True

    def evaluate(self):
        self.train = False

    def forward(self, inputs):
        q, i = (np.ascontiguousarray(a, dtype=np.float32) for a in inputs)
        m = self._model
        m.set_params(BLOCK_MULTIMODAL, self._w)
        n = q.shape[0]
        Q_, I_ = DeviceBuffer(m, q), DeviceBuffer(m, i)
        O_ = DeviceBuffer(m, np.zeros((n, m.cfg.C), dtype=np.float32))
        mq = mi = None
        if self.train:
            if self.masks is None:
                raise _lib.NvqaError("training-mode forward needs explicit .masks")
            mq, mi = (DeviceBuffer(m, np.ascontiguousarray(a, dtype=np.float32)) for a in self.masks)
        _lib.check(m.lib.nvqa_axb_forward(m.handle, Q_.ptr, I_.ptr, None if mq is None else mq.ptr,
                                          None if mi is None else mi.ptr, n, O_.ptr))
        self.output = O_.get()
        return self.output


class netdef:
    @staticmethod
    def AxB(nhA, nhB, nhcommon, dropout=0, rnn_size=512, rnn_layers=2):
        return _AxB(nhA, nhB, nhcommon, dropout, rnn_size, rnn_layers)


# ---- nn.CrossEntropyCriterion -------------------------------------------------------------------------------------
class CrossEntropyCriterion:
    """criterion:forward(scores, labels) -> number; criterion:backward(scores, labels) -> dscores (sizeAverage)."""

    def __init__(self, num_output):
        self._model = Arch1Model(Arch1Config(E=4, H=4, L=1, V=8, I=4, C=4, O=num_output, T=2, B=1024))
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
