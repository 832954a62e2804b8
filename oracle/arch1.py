"""NumPy restatement of the reference's arch1 VQA training / eval step (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py ("parity unpinned").  All citations are
relative to /root/reference/002_train_vqa_arch1/.  The control flow follows ``JdJ``
(002_train_baseline.lua:272-335) op for op, in the reference's own *packed, length-sorted*
layout (misc/RNNUtils.lua:84-211), so that the CUDA path -- which uses a padded time-major
layout with per-row activity masks -- is checked against the reference's formulation and not
against a copy of itself.

dtype is a parameter: ``np.float64`` is the "truth" used for finite-difference checks,
``np.float32`` mirrors what Torch7's FloatTensor path computes.
"""
from dataclasses import dataclass

import numpy as np

from . import rng


@dataclass
class Arch1Config:
    V: int = 14773      # vocabulary size          002_train_baseline.lua:126-127 (vocab_oracle.json)
    E: int = 200        # input_encoding_size      002_train_baseline.lua:34
    H: int = 512        # rnn_size                 :35
    L: int = 2          # rnn_layer                :36
    I: int = 4096       # nhimage                  :33
    C: int = 1024       # common_embedding_size    :37
    O: int = 1000       # num_output               :38
    T: int = 26         # buffer_size_q = question matrix width   :136
    p: float = 0.5      # every Dropout on the path uses 0.5      :143,147,152,153
    fusion_skip: bool = False   # netdef.AskipB instead of netdef.AxB (003_train_ae_based_wp.lua:151)

    @property
    def S(self):        # width of the packed LSTM state [c1 h1 c2 h2 ...]  misc/LSTM.lua:21-22,70
        return 2 * self.L * self.H

    def enc_layout(self):
        """(name, shape) of the encoder's flat parameter vector, nngraph order:
        per layer i2h.weight, i2h.bias, h2h.weight, h2h.bias  (misc/LSTM.lua:41-42)."""
        out = []
        for l in range(self.L):
            n_in = self.E if l == 0 else self.H
            out += [(f"Wi{l}", (4 * self.H, n_in)), (f"bi{l}", (4 * self.H,)),
                    (f"Wh{l}", (4 * self.H, self.H)), (f"bh{l}", (4 * self.H,))]
        return out

    def emb_layout(self):
        """nn.Linear(V,E): weight[E x V] then bias[E]   002_train_baseline.lua:142,174."""
        return [("We", (self.E, self.V)), ("be", (self.E,))]

    def mm_layout(self):
        """AxB(q: 2LH -> C, i: I -> C) then Linear(C,O)   misc/netdef.lua:6-14, 002_train_baseline.lua:151-154."""
        return [("Wq", (self.C, self.S)), ("bq", (self.C,)), ("Wv", (self.C, self.I)), ("bv", (self.C,)),
                ("Wc", (self.O, self.C)), ("bc", (self.O,))]

    @staticmethod
    def _size(layout):
        return int(sum(int(np.prod(s)) for _, s in layout))

    @property
    def n_enc(self):
        return self._size(self.enc_layout())

    @property
    def n_emb(self):
        return self._size(self.emb_layout())

    @property
    def n_mm(self):
        return self._size(self.mm_layout())


def split_flat(w, layout):
    """Views of a flat vector by (name, shape) -- split_vector (misc/RNNUtils.lua:25-39)."""
    out, off = {}, 0
    for name, shape in layout:
        n = int(np.prod(shape))
        out[name] = w[off:off + n].reshape(shape)
        off += n
    assert off == w.size
    return out


# ----------------------------------------------------------------------------------------------
# integer preprocessing  (bit-exact part of the path)
# ----------------------------------------------------------------------------------------------
def right_align(seq, lengths):
    """misc/RNNUtils.lua:54-61: shift every question to the right edge, zero-pad on the left."""
    seq = np.asarray(seq)
    v = np.zeros_like(seq)
    N = seq.shape[1]
    for i in range(seq.shape[0]):
        n = int(lengths[i])
        v[i, N - n:N] = seq[i, :n]
    return v


def inverse_mapping(ind):
    """misc/RNNUtils.lua:13-16: y = x[ind]  ->  x = y[inv]."""
    return np.argsort(ind, kind="stable")


def sort_encoding_right_align(q_ra, lengths):
    """misc/RNNUtils.lua:105-124 without the one-hot expansion (the one-hot Linear is a gather).

    Returns (words[N], batch_sizes[L], sort_index[B], sort_index_inverse[B]); permutations are
    0-based here (Torch's are 1-based).  torch.sort(.., true) is an unstable quicksort whose tie
    order is unknowable; the restatement uses a *stable* descending sort.
    """
    q_ra = np.asarray(q_ra)
    lengths = np.asarray(lengths).astype(np.int64)
    sort_index = np.argsort(-lengths, kind="stable")
    len_sorted = lengths[sort_index]
    inv = inverse_mapping(sort_index)
    D = q_ra.shape[1]
    L = int(len_sorted[0])
    qt = q_ra[sort_index].T[D - L:D]                 # [L x B]
    words, batch_sizes = [], np.zeros(L, dtype=np.int64)
    for i in range(L):
        n = int(np.sum(len_sorted >= L - i))         # ge(L-i+1) with 1-based i
        words.append(qt[i, :n])
        batch_sizes[i] = n
    words = np.concatenate(words) if words else np.zeros(0, dtype=q_ra.dtype)
    return words.astype(np.int64), batch_sizes, sort_index.astype(np.int64), inv.astype(np.int64)


def l2_normalize_rows(x, split=0):
    """002_train_baseline.lua:117-123: x / sqrt(sum(x*x, 2)), no epsilon.  split > 0: the column blocks [0, split) and
    [split, I) are normalised separately (early fusion, 003_train_ae_based_ef.lua:116-124)."""
    if split:
        return np.concatenate([l2_normalize_rows(x[:, :split]), l2_normalize_rows(x[:, split:])], axis=1)
    nm = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
    return (x / nm).astype(x.dtype)


# ----------------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def embedding_forward(emb, words, mask):
    """nn.Linear(V,E) on a one-hot row == column gather + bias; then Dropout, Tanh
    (002_train_baseline.lua:141-144,300).  ``words`` are 1-based token ids."""
    pre = emb["We"][:, words - 1].T + emb["be"]
    if mask is not None:
        pre = pre * mask
    return np.tanh(pre)


def embedding_backward(emb_grad, words, y, dy, mask):
    """Tanh / Dropout backward then the one-hot Linear's accGradParameters (002_train_baseline.lua:320)."""
    dpre = dy * (1.0 - y * y)
    if mask is not None:
        dpre = dpre * mask
    np.add.at(emb_grad["We"].T, words - 1, dpre)      # dW[:, w] += dpre
    emb_grad["be"] += dpre.sum(axis=0)


def lstm_cell_forward(cfg, enc, state, x, masks):
    """LSTM.lstm_conventional forward (misc/LSTM.lua:12-73).  state = [c1 h1 c2 h2 ...];
    gate rows in the order i, f, o, g (:45-52); layer L>1 input = Dropout(h'_{L-1}) (:36-37)."""
    H = cfg.H
    outs, cache = [], []
    inp = x
    for l in range(cfg.L):
        c_prev = state[:, 2 * l * H:(2 * l + 1) * H]
        h_prev = state[:, (2 * l + 1) * H:(2 * l + 2) * H]
        if l > 0:
            inp = outs[-1]
            if masks is not None:
                inp = inp * masks[l - 1]
        a = inp @ enc[f"Wi{l}"].T + enc[f"bi{l}"] + h_prev @ enc[f"Wh{l}"].T + enc[f"bh{l}"]
        sg = _sigmoid(a[:, :3 * H])
        i, f, o = sg[:, :H], sg[:, H:2 * H], sg[:, 2 * H:3 * H]
        g = np.tanh(a[:, 3 * H:])
        c = f * c_prev + i * g
        tc = np.tanh(c)
        h = o * tc
        outs += [c, h]
        cache.append((inp, c_prev, h_prev, i, f, o, g, tc))
    return np.concatenate(outs, axis=1), cache


def lstm_cell_backward(cfg, enc, enc_grad, cache, dstate, masks):
    """gModule backward of the cell: top layer first; gradients of h'_l are summed
    over its two consumers (output state and layer l+1); both biases receive sum(da)."""
    H = cfg.H
    n = dstate.shape[0]
    dprev = np.zeros((n, cfg.S), dtype=dstate.dtype)
    dh_from_above = None
    dx = None
    for l in reversed(range(cfg.L)):
        inp, c_prev, h_prev, i, f, o, g, tc = cache[l]
        dc = dstate[:, 2 * l * H:(2 * l + 1) * H]
        dh = dstate[:, (2 * l + 1) * H:(2 * l + 2) * H]
        if dh_from_above is not None:
            dh = dh + dh_from_above
        dc = dc + dh * o * (1.0 - tc * tc)
        da = np.concatenate([dc * g * i * (1.0 - i), dc * c_prev * f * (1.0 - f),
                             dh * tc * o * (1.0 - o), dc * i * (1.0 - g * g)], axis=1)
        enc_grad[f"Wi{l}"] += da.T @ inp
        enc_grad[f"bi{l}"] += da.sum(axis=0)
        enc_grad[f"Wh{l}"] += da.T @ h_prev
        enc_grad[f"bh{l}"] += da.sum(axis=0)
        dprev[:, 2 * l * H:(2 * l + 1) * H] = dc * f
        dprev[:, (2 * l + 1) * H:(2 * l + 2) * H] = da @ enc[f"Wh{l}"]
        dinp = da @ enc[f"Wi{l}"]
        if l > 0:
            if masks is not None:
                dinp = dinp * masks[l - 1]
            dh_from_above = dinp
        else:
            dx = dinp
    return dprev, dx


def rnn_forward(cfg, enc, init_state, inputs, sizes, lstm_masks):
    """misc/RNNUtils.lua:128-154 for right-aligned (non-decreasing) ``sizes``.  Newly active rows
    start from ``init_state`` (zeros); the reference's view-aliasing bug (App. C-2) is NOT
    replicated.  Returns the T+1 states (as fed to each step, grown to that step's size) + caches."""
    N = len(sizes)
    states = [init_state[:sizes[0]].copy()]
    caches = []
    for i in range(N):
        if i > 0:
            assert sizes[i] >= sizes[i - 1], "oracle covers the right-aligned (growing) case only"
            if sizes[i] > sizes[i - 1]:
                pad = init_state[:sizes[i]].copy()
                pad[:sizes[i - 1]] = states[i]
                states[i] = pad
        nxt, cache = lstm_cell_forward(cfg, enc, states[i], inputs[i],
                                       None if lstm_masks is None else lstm_masks[i])
        states.append(nxt)
        caches.append(cache)
    return states, caches


def rnn_backward(cfg, enc, enc_grad, dend_state, caches, sizes, lstm_masks):
    """misc/RNNUtils.lua:181-210 (the branch taken when doutputs is not a table)."""
    N = len(sizes)
    dstate = dend_state[:sizes[N - 1]]
    dinputs = [None] * N
    for i in reversed(range(N)):
        dprev, dx = lstm_cell_backward(cfg, enc, enc_grad, caches[i], dstate,
                                       None if lstm_masks is None else lstm_masks[i])
        dinputs[i] = dx
        dstate = dprev if (i == 0 or sizes[i] == sizes[i - 1]) else dprev[:sizes[i - 1]]
    return dstate, dinputs


def multimodal_forward(cfg, mm, tv_q, fv_im, masks):
    """netdef.AxB (misc/netdef.lua:6-14) -> Dropout -> Linear(C,O) (002_train_baseline.lua:151-154)."""
    qd = tv_q if masks is None else tv_q * masks["q"]
    vd = fv_im if masks is None else fv_im * masks["i"]
    qc = np.tanh(qd @ mm["Wq"].T + mm["bq"])
    ic = np.tanh(vd @ mm["Wv"].T + mm["bv"])
    z = qc * ic
    if getattr(cfg, "fusion_skip", False):      # netdef.AskipB (misc/netdef.lua:16-25): CAddTable({qc, qc * ic})
        z = qc + z
    zd = z if masks is None else z * masks["z"]
    scores = zd @ mm["Wc"].T + mm["bc"]
    return scores, (qd, vd, qc, ic, zd)


def multimodal_backward(cfg, mm, mm_grad, cache, dscores, masks):
    qd, vd, qc, ic, zd = cache
    mm_grad["Wc"] += dscores.T @ zd
    mm_grad["bc"] += dscores.sum(axis=0)
    dz = dscores @ mm["Wc"]
    if masks is not None:
        dz = dz * masks["z"]
    dqc = dz * ic
    if getattr(cfg, "fusion_skip", False):
        dqc = dqc + dz
    dqpre = dqc * (1.0 - qc * qc)
    dipre = dz * qc * (1.0 - ic * ic)
    mm_grad["Wq"] += dqpre.T @ qd
    mm_grad["bq"] += dqpre.sum(axis=0)
    mm_grad["Wv"] += dipre.T @ vd
    mm_grad["bv"] += dipre.sum(axis=0)
    dq = dqpre @ mm["Wq"]
    if masks is not None:
        dq = dq * masks["q"]
    return dq          # d fv_im is computed by the reference but never used


def cross_entropy(scores, labels):
    """nn.CrossEntropyCriterion = LogSoftMax + ClassNLLCriterion(sizeAverage) (002_train_baseline.lua:157,308-310).
    labels are 1-based.  Returns (f, dscores)."""
    B = scores.shape[0]
    mx = scores.max(axis=1, keepdims=True)
    ex = np.exp(scores - mx)
    se = ex.sum(axis=1, keepdims=True)
    logp = scores - mx - np.log(se)
    f = -logp[np.arange(B), labels - 1].sum() / B
    d = ex / se
    d[np.arange(B), labels - 1] -= 1.0
    return scores.dtype.type(f), (d / B).astype(scores.dtype)


def argmax_first(scores):
    """torch.max(scores, 2): first index attaining the maximum, 1-based (004_eval_model.lua:233)."""
    return (np.argmax(scores, axis=1) + 1).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# dropout masks in the reference's packed layout
# ----------------------------------------------------------------------------------------------
def build_masks(cfg, seed, B, sizes, sort_index, dtype=np.float32):
    """Masks for every Dropout site from the shared counter hash (oracle/rng.py).  Element indices are
    defined on the *padded* geometry (absolute time t, ORIGINAL batch row b, feature j) so that the
    packed oracle and the padded CUDA path draw identical masks."""
    T, E, H = cfg.T, cfg.E, cfg.H
    Lq = len(sizes)
    emb, lstm = [], []
    for i in range(Lq):
        t = T - Lq + i
        rows = sort_index[:sizes[i]].astype(np.int64)
        base = (t * B + rows)[:, None]
        emb.append(rng.keep_scale(seed, rng.STREAM_EMB, base * E + np.arange(E)[None, :], cfg.p, dtype))
        lstm.append([rng.keep_scale(seed, rng.STREAM_LSTM0 + l, base * H + np.arange(H)[None, :], cfg.p, dtype)
                     for l in range(cfg.L - 1)])
    b = np.arange(B, dtype=np.int64)[:, None]
    return {
        "emb": np.concatenate(emb, axis=0),
        "lstm": lstm,
        "q": rng.keep_scale(seed, rng.STREAM_AXB_Q, b * cfg.S + np.arange(cfg.S)[None, :], cfg.p, dtype),
        "i": rng.keep_scale(seed, rng.STREAM_AXB_I, b * cfg.I + np.arange(cfg.I)[None, :], cfg.p, dtype),
        "z": rng.keep_scale(seed, rng.STREAM_HEAD, b * cfg.C + np.arange(cfg.C)[None, :], cfg.p, dtype),
    }


def pack_masks(cfg, padded, B, sizes, sort_index):
    """Padded-layout masks {emb [T,B,E], lstm [L-1,T,B,H], q, i, z} -> the packed layout used above."""
    Lq = len(sizes)
    emb, lstm = [], []
    for i in range(Lq):
        t = cfg.T - Lq + i
        rows = sort_index[:sizes[i]]
        emb.append(padded["emb"][t, rows])
        lstm.append([padded["lstm"][l, t, rows] for l in range(cfg.L - 1)])
    return {"emb": np.concatenate(emb, axis=0), "lstm": lstm, "q": padded["q"], "i": padded["i"], "z": padded["z"]}


# ----------------------------------------------------------------------------------------------
# the step
# ----------------------------------------------------------------------------------------------
def forward(cfg, enc_w, emb_w, mm_w, q_ra, lengths, fv_im, seed=None, dtype=np.float32, masks=None):
    """Forward of JdJ / 004_eval_model.lua:202-218.  ``seed=None`` and ``masks=None`` -> evaluate mode
    (Dropout = identity); ``masks`` (packed layout, see build_masks) overrides the hash generator.
    ``fv_im`` must already be L2-normalised (the reference does that at load time, :117-123)."""
    enc = split_flat(enc_w.astype(dtype), cfg.enc_layout())
    emb = split_flat(emb_w.astype(dtype), cfg.emb_layout())
    mm = split_flat(mm_w.astype(dtype), cfg.mm_layout())
    B = q_ra.shape[0]
    words, sizes, sort_index, inv = sort_encoding_right_align(q_ra, lengths)
    if masks is None and seed is not None:
        masks = build_masks(cfg, seed, B, sizes, sort_index, dtype)
    y = embedding_forward(emb, words, None if masks is None else masks["emb"])
    offs = np.concatenate([[0], np.cumsum(sizes)])
    inputs = [y[offs[i]:offs[i + 1]] for i in range(len(sizes))]          # split_vector  :300
    init = np.zeros((B, cfg.S), dtype=dtype)                               # repeatTensor(dummy_state) :303
    states, caches = rnn_forward(cfg, enc, init, inputs, sizes, None if masks is None else masks["lstm"])
    tv_q = states[-1][inv]                                                 # :306
    scores, mcache = multimodal_forward(cfg, mm, tv_q, fv_im.astype(dtype), masks)
    ctx = dict(enc=enc, emb=emb, mm=mm, words=words, sizes=sizes, sort_index=sort_index, inv=inv, masks=masks,
               y=y, offs=offs, caches=caches, mcache=mcache, tv_q=tv_q)
    return scores, ctx


def jdj(cfg, enc_w, emb_w, mm_w, q_ra, lengths, fv_im, labels, seed=None, dtype=np.float32,
        grad_scale=1.0, clamp=10.0, masks=None, lr_scale=1.0):
    """f, gradients of 002_train_baseline.lua:272-335.  Gradients are returned as the three flat
    blocks in the optimiser's order (encoder, embedding, multimodal) (:328), clamped to +-10 (:329).
    ``grad_scale`` (default 1) is the 1/n_ranks factor of the data-parallel extension, applied
    before the clamp."""
    scores, ctx = forward(cfg, enc_w, emb_w, mm_w, q_ra, lengths, fv_im, seed, dtype, masks)
    masks = ctx["masks"]
    f, dscores = cross_entropy(scores, np.asarray(labels).astype(np.int64))   # :308-310
    g_enc = np.zeros(cfg.n_enc, dtype=dtype)
    g_emb = np.zeros(cfg.n_emb, dtype=dtype)
    g_mm = np.zeros(cfg.n_mm, dtype=dtype)
    enc_grad = split_flat(g_enc, cfg.enc_layout())
    emb_grad = split_flat(g_emb, cfg.emb_layout())
    mm_grad = split_flat(g_mm, cfg.mm_layout())
    dtv = multimodal_backward(cfg, ctx["mm"], mm_grad, ctx["mcache"], dscores, masks)   # :312
    dend = dtv[ctx["sort_index"]]                                                       # :313
    _, dinputs = rnn_backward(cfg, ctx["enc"], enc_grad, dend, ctx["caches"], ctx["sizes"],
                              None if masks is None else masks["lstm"])                 # :316
    dy = np.concatenate(dinputs, axis=0)                                                # join_vector :319
    embedding_backward(emb_grad, ctx["words"], ctx["y"], dy, None if masks is None else masks["emb"])  # :320
    grads = [g_enc, g_emb, g_mm]
    if lr_scale != 1.0:     # join_vector({encoder_adw_q * lr_scale, embedding_dw_q * lr_scale, multimodal_dw})  003_train_ae_based_wp.lua:344
        grads = [g_enc * dtype(lr_scale), g_emb * dtype(lr_scale), g_mm]
    if grad_scale != 1.0:
        grads = [g * dtype(grad_scale) for g in grads]
    if clamp is not None:
        grads = [np.clip(g, -clamp, clamp) for g in grads]
    return f, grads, scores, ctx


def rmsprop_update(x, g, m, lr, alpha=0.99, eps=1e-8, wd=0.0):
    """optim.rmsprop as restated in-repo by misc/rmsprop_lrscale.lua:14-34 (lrs == 1):
    g += wd*x;  m = alpha*m + (1-alpha)*g*g;  x -= lr * g / (sqrt(m) + eps).  In place."""
    dt = x.dtype.type
    if wd != 0.0:
        g = g + dt(wd) * x
    m *= dt(alpha)
    m += dt(1.0 - alpha) * g * g
    x -= dt(lr) * (g / (np.sqrt(m) + dt(eps)))
    return x, m


def train_step(cfg, enc_w, emb_w, mm_w, m_state, batch, lr, seed=None, dtype=np.float32, grad_scale=1.0):
    """One iteration of the training loop (002_train_baseline.lua:408-410) on explicit inputs.
    The optimiser vector is join{encoder, embedding, multimodal} (:183,190).  Updates in place and
    returns (f, lr_next)."""
    q_ra, lengths, fv_im, labels = batch
    f, grads, _, _ = jdj(cfg, enc_w, emb_w, mm_w, q_ra, lengths, fv_im, labels, seed, dtype, grad_scale)
    for w, g, m in zip((enc_w, emb_w, mm_w), grads, m_state):
        rmsprop_update(w, g.astype(w.dtype), m, lr)
    return f, lr * 0.99997592083        # decay_factor :78,410


def multiple_choice_select(scores_row, mc_ids):
    """004_eval_model.lua:257-271: argmax of scores restricted to the non-zero candidate ids (1-based)."""
    ids = [int(a) for a in mc_ids if a != 0]
    vals = np.array([scores_row[a - 1] for a in ids])
    return ids[int(np.argmax(vals))]
