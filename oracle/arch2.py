"""NumPy restatement of the reference's arch2 VQA training / eval step (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py ("parity unpinned").  Citations are relative to
/root/reference/003_train_vqa_arch2/.  Follows ``JdJ`` (002_train_baseline.lua:277-333) and ``nn.Encoder``
(misc/Encoder_lstm.lua:152-263) with the cell of misc/LSTM_encoder.lua:5-57 (same gate order i,f,o,g and the same
Linear order as arch1's cell, so oracle.arch1's cell functions are reused).

Sequence fed to the LSTM (Encoder_lstm.lua:170-221): step 1 = image projected by Linear(I,E), step 2 = START token
(id V+1), steps 3.. = the question words, NOT right-aligned (002_train_baseline.lua:216); zeros are replaced by
token 1 and processed unmasked (:197); a time row that is all zeros is skipped, so the number of
executed steps is tmax = 2 + (longest question in the batch).  Output = top-layer h at tmax (:224).
Two literal-reference behaviours are flags (default off = the evident intent):
``literal_lookup_grad`` -- createClones builds the per-step lookup tables from ``self.lookup_table:clone('weight')``
(Encoder_lstm.lua:53), which shares no gradWeight with the module whose gradient parameters() returns, so the
LookupTable block of the gradient stays zero (it only sees weight decay); ``h0_top`` -- updateGradInput stores the
head's gradInput tensor into ``self.init_state_enc[num_state]`` (:238-239) and _createInitState re-zeroes it only on a
batch-size change (:37-40), so from the second training step on the top layer starts from h0 = the PREVIOUS step's
d loss / d h_T.  ``train_step(..., literal=True, carry=...)`` threads that tensor through.
"""
from dataclasses import dataclass

import numpy as np

from . import arch1 as A
from . import rng


@dataclass
class Arch2Config:
    V: int = 14773      # vocabulary size
    E: int = 512        # -input_encoding_size   002_train_baseline.lua:37
    H: int = 512        # -rnn_size              :38
    L: int = 1          # -num_layers            :39
    I: int = 4096       # -nhimage               :43  (2048 for the Inception features of BASELINE config 4)
    O: int = 1000       # -num_output            :41
    T: int = 26         # seq_length = question matrix width
    p: float = 0.5      # -drop_prob_ae (between LSTM layers) and the head Dropout(0.5) :162

    @property
    def S(self):
        return 2 * self.L * self.H

    def cnn_layout(self):
        """cnn_projection = nn.Linear(nhimage, input_encoding_size)   :166"""
        return [("Wcnn", (self.E, self.I)), ("bcnn", (self.E,))]

    def enc_layout(self):
        """nn.Encoder:parameters(): the LSTM core, then LookupTable(V+1, E)   misc/Encoder_lstm.lua:66-83"""
        out = []
        for l in range(self.L):
            n_in = self.E if l == 0 else self.H
            out += [(f"Wi{l}", (4 * self.H, n_in)), (f"bi{l}", (4 * self.H,)),
                    (f"Wh{l}", (4 * self.H, self.H)), (f"bh{l}", (4 * self.H,))]
        return out + [("lookup", (self.V + 1, self.E))]

    def mm_layout(self):
        """multimodal_net = Dropout(0.5) -> Linear(rnn_size, noutput)   :162-164"""
        return [("Wc", (self.O, self.H)), ("bc", (self.O,))]

    n_cnn = property(lambda self: A.Arch1Config._size(self.cnn_layout()))
    n_enc = property(lambda self: A.Arch1Config._size(self.enc_layout()))
    n_mm = property(lambda self: A.Arch1Config._size(self.mm_layout()))


def step_tokens(cfg, seq):
    """Token ids fed at steps 2.. (Encoder_lstm.lua:177-203) and the number of executed steps tmax."""
    seq = np.asarray(seq)
    B = seq.shape[0]
    toks = [None, np.full(B, cfg.V + 1, dtype=np.int64)]          # step 1 is the image, step 2 the START token
    tmax = 2
    for k in range(cfg.T):
        it = seq[:, k].astype(np.int64).copy()
        if it.sum() == 0:                                           # can_skip: all sequences have terminated
            toks.append(None)
            continue
        it[it == 0] = 1
        toks.append(it)
        tmax = k + 3
    return toks, tmax


def build_masks(cfg, seed, B, steps, dtype=np.float32):
    b = np.arange(B, dtype=np.int64)[:, None]
    lstm = []
    for s in range(steps):
        base = (s * B + b)
        lstm.append([rng.keep_scale(seed, rng.STREAM_LSTM0 + l, base * cfg.H + np.arange(cfg.H)[None, :], cfg.p, dtype)
                     for l in range(cfg.L - 1)])
    return {"lstm": lstm, "z": rng.keep_scale(seed, rng.STREAM_HEAD, b * cfg.H + np.arange(cfg.H)[None, :], cfg.p, dtype)}


def jdj(cfg, cnn_w, enc_w, mm_w, seq, fv_im, labels, seed=None, dtype=np.float32, clamp=10.0, grad_scale=1.0,
        literal_lookup_grad=False, h0_top=None):
    """f, gradients (cnn, encoder, multimodal -- the optimiser's order, :192,326), scores, context.
    ctx["dz"] is the head's gradInput (what the literal reference leaves in init_state_enc[num_state])."""
    cnn = A.split_flat(cnn_w.astype(dtype), cfg.cnn_layout())
    enc = A.split_flat(enc_w.astype(dtype), cfg.enc_layout())
    mm = A.split_flat(mm_w.astype(dtype), cfg.mm_layout())
    B = seq.shape[0]
    toks, tmax = step_tokens(cfg, seq)
    masks = None if seed is None else build_masks(cfg, seed, B, tmax, dtype)
    fv = fv_im.astype(dtype)
    x0 = fv @ cnn["Wcnn"].T + cnn["bcnn"]                           # cnn_projection:forward   :308
    state = np.zeros((B, cfg.S), dtype=dtype)
    if h0_top is not None:                                          # literal reference, 2nd step on (App. C-5)
        state[:, (2 * cfg.L - 1) * cfg.H:2 * cfg.L * cfg.H] = h0_top
    caches, xs = [], []
    for s in range(tmax):
        x = x0 if s == 0 else enc["lookup"][toks[s] - 1]
        xs.append(x)
        state, cache = A.lstm_cell_forward(cfg, enc, state, x, None if masks is None else masks["lstm"][s])
        caches.append(cache)
    H = cfg.H
    out = state[:, (2 * cfg.L - 1) * H:2 * cfg.L * H]               # top-layer h at tmax   Encoder_lstm.lua:224
    zd = out if masks is None else out * masks["z"]
    scores = zd @ mm["Wc"].T + mm["bc"]
    g_cnn = np.zeros(cfg.n_cnn, dtype=dtype)
    g_enc = np.zeros(cfg.n_enc, dtype=dtype)
    g_mm = np.zeros(cfg.n_mm, dtype=dtype)
    if labels is None:
        return None, None, scores, dict(out=out, tmax=tmax)
    f, dscores = A.cross_entropy(scores, np.asarray(labels).astype(np.int64))
    cnn_g, enc_g, mm_g = (A.split_flat(g_cnn, cfg.cnn_layout()), A.split_flat(g_enc, cfg.enc_layout()),
                          A.split_flat(g_mm, cfg.mm_layout()))
    mm_g["Wc"] += dscores.T @ zd
    mm_g["bc"] += dscores.sum(axis=0)
    dz = dscores @ mm["Wc"]
    if masks is not None:
        dz = dz * masks["z"]
    dstate = np.zeros((B, cfg.S), dtype=dtype)
    dstate[:, (2 * cfg.L - 1) * H:2 * cfg.L * H] = dz              # Encoder_lstm.lua:238-239
    for s in reversed(range(tmax)):
        dstate, dx = A.lstm_cell_backward(cfg, enc, enc_g, caches[s], dstate, None if masks is None else masks["lstm"][s])
        if s == 0:
            cnn_g["Wcnn"] += dx.T @ fv                              # cnn_projection:backward   :322
            cnn_g["bcnn"] += dx.sum(axis=0)
        else:
            np.add.at(enc_g["lookup"], toks[s] - 1, dx)             # LookupTable accGradParameters   Encoder_lstm.lua:256
    if literal_lookup_grad:                                         # the clones' gradWeight is not the one parameters() returns
        enc_g["lookup"][...] = 0
    grads = [g_cnn, g_enc, g_mm]
    if grad_scale != 1.0:
        grads = [g * dtype(grad_scale) for g in grads]
    if clamp is not None:
        grads = [np.clip(g, -clamp, clamp) for g in grads]
    return f, grads, scores, dict(out=out, tmax=tmax, dz=dz)


def train_step(cfg, cnn_w, enc_w, mm_w, m_state, batch, lr, seed=None, dtype=np.float32, wd=1e-4, literal=False, carry=None):
    """optim.rmsprop(JdJ, ...) with optimize.weightDecay = 1e-4 (:197): wd*x is added AFTER the clamp.
    literal=True: both literal-reference behaviours; ``carry`` = {} kept by the caller across steps (holds the stale h0
    and the batch size it belongs to)."""
    seq, fv_im, labels = batch
    h0 = None
    if literal and carry is not None and carry.get("B") == seq.shape[0]:
        h0 = carry.get("dz")
    f, grads, _, ctx = jdj(cfg, cnn_w, enc_w, mm_w, seq, fv_im, labels, seed, dtype, literal_lookup_grad=literal, h0_top=h0)
    if literal and carry is not None:
        carry["dz"], carry["B"] = ctx["dz"], seq.shape[0]
    for w, g, m in zip((cnn_w, enc_w, mm_w), grads, m_state):
        A.rmsprop_update(w, g.astype(w.dtype), m, lr, wd=wd)
    return f, lr * 0.99997592083
