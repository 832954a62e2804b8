"""NumPy restatement of the reference's arch1 text autoencoder training step (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py ("parity unpinned").  Citations are relative to
/root/reference/001_train_autoencoder/ unless they name another stage directory.

Follows ``lossFun`` (001_train_arch1_text_autoencoder.lua:208-249), ``nn.AutoEncoder``
(misc/AutoEncoder_text_nostart.lua:222-393) and ``nn.LanguageModelCriterion`` (:414-455):

* lookup_table = LookupTable(V+1, E) -> Dropout(0.5) -> Tanh (:29-32), shared by encoder and decoder, one Dropout
  mask per timestep clone (:68-82).
* encoder = LSTM_encoder.lstm (misc/LSTM_encoder.lua:5-57): steps t = 1..tmax over seq[t] with zeros replaced by token 1
  and processed UNMASKED (:258-266); an all-zero time row ends the loop (tmax, :249-256,281); initial state zeros.
* decoder = the cell of 003_train_vqa_arch2/misc/LSTM_decoder.lua:5-64: same LSTM core, then
  Dropout(p)(top_h) -> Linear(H, V+1) -> LogSoftMax.  Step 1 is fed START (= V+1), step t >= 2 is fed seq[t-1]
  (zeros -> token 1); initial state = encoder state at tmax (:287-288); steps 1..tmax+1 are executed (:296-337).
* criterion: target of decoder step t is seq[t] (t <= D), the first null becomes END (= V+1), later nulls are
  skipped; loss = -sum(logp[target]) / n, gradient -1/n at the targets (:427-449).
* backward: decoder steps tmax+1..1 starting from zero state gradients, then encoder steps tmax..1 starting from the
  decoder's d(initial state) (:345-387); gradients clamped to +-grad_clip (0.1), then += weight_decay (1e-6) * params
  (001_train_arch1_text_autoencoder.lua:237-243); Adam of misc/optim_updates.lua:78-111 (eps outside the sqrt).

Reference quirk kept switchable (``literal_lookup_grad``): createClones builds the per-step lookup tables from
``self.lookup_table:clone('weight')`` (:64-66), which shares the weight but NOT gradWeight with the module whose
gradWeight ``parameters()`` returns (:94,100), so in the literal reference the LookupTable block of grad_params stays
zero and the table only sees the weight-decay term.  The intended semantics (gradient accumulated over all encoder
and decoder steps) is the default here and in the CUDA path; the literal behaviour is
``literal_lookup_grad=True``.
"""
from dataclasses import dataclass

import numpy as np

from . import arch1 as A
from . import rng

STREAM_AE_ENC_EMB = 32    # lookup Dropout(0.5) of the encoder clones      AutoEncoder_text_nostart.lua:31,66
STREAM_AE_DEC_EMB = 33    # lookup Dropout(0.5) of the decoder clones      :67
STREAM_AE_OUT = 34        # 'drop_final' on the decoder's top h            003_train_vqa_arch2/misc/LSTM_decoder.lua:56


@dataclass
class AEConfig:
    V: int = 20000      # vocab_size (000_prepro_book_corpus.py:270); tables have V+1 rows (START/END = V+1)
    E: int = 512        # -input_encoding_size   001_train_arch1_text_autoencoder.lua:29
    H: int = 512        # -rnn_size              :28
    L: int = 1          # -num_layers            :30
    T: int = 16         # seq_length             000_prepro_book_corpus.py:265
    p: float = 0.5      # -drop_prob_ae          :36 (the lookup Dropout is a hard-coded 0.5)

    @property
    def S(self):
        return 2 * self.L * self.H

    def core_layout(self):
        out = []
        for l in range(self.L):
            n_in = self.E if l == 0 else self.H
            out += [(f"Wi{l}", (4 * self.H, n_in)), (f"bi{l}", (4 * self.H,)),
                    (f"Wh{l}", (4 * self.H, self.H)), (f"bh{l}", (4 * self.H,))]
        return out

    def enc_layout(self):
        """self.encoder:parameters()   AutoEncoder_text_nostart.lua:88"""
        return self.core_layout()

    def dec_layout(self):
        """self.decoder:parameters(): LSTM core, then the 'decoder' Linear(H, V+1)   :89"""
        return self.core_layout() + [("Wd", (self.V + 1, self.H)), ("bd", (self.V + 1,))]

    def lut_layout(self):
        """self.lookup_table:parameters()   :90"""
        return [("lookup", (self.V + 1, self.E))]

    n_enc = property(lambda self: A.Arch1Config._size(self.enc_layout()))
    n_dec = property(lambda self: A.Arch1Config._size(self.dec_layout()))
    n_lut = property(lambda self: A.Arch1Config._size(self.lut_layout()))


def step_tokens(cfg, seq):
    """seq [B x T] (zero-padded).  Returns encoder tokens per step, decoder tokens per step, tmax.
    AutoEncoder_text_nostart.lua:244-282 (encoder) and :296-337 (decoder)."""
    seq = np.asarray(seq).astype(np.int64)
    B = seq.shape[0]
    enc, tmax = [], 0
    for t in range(cfg.T):
        it = seq[:, t].copy()
        if it.sum() == 0:
            enc.append(None)
            continue
        it[it == 0] = 1
        enc.append(it)
        tmax = t + 1
    dec = [np.full(B, cfg.V + 1, dtype=np.int64)]
    for t in range(1, cfg.T + 1):
        it = seq[:, t - 1].copy()
        if it.sum() == 0:
            dec.append(None)
            continue
        it[it == 0] = 1
        dec.append(it)
    return enc, dec, tmax


def lm_targets(cfg, seq):
    """nn.LanguageModelCriterion target selection (:427-447).  Returns targets [T+1 x B] (0 = no prediction) and n."""
    seq = np.asarray(seq).astype(np.int64)
    B = seq.shape[0]
    L = cfg.T + 1
    tg = np.zeros((L, B), dtype=np.int64)
    for b in range(B):
        first = True
        for t in range(L):
            ti = seq[b, t] if t < cfg.T else 0
            if ti == 0 and first:
                ti = cfg.V + 1
                first = False
            tg[t, b] = ti
    return tg, int((tg != 0).sum())


def build_masks(cfg, seed, B, tmax, dtype=np.float32):
    b = np.arange(B, dtype=np.int64)[:, None]
    e = np.arange(cfg.E, dtype=np.int64)[None, :]
    h = np.arange(cfg.H, dtype=np.int64)[None, :]
    return {
        "enc_emb": [rng.keep_scale(seed, STREAM_AE_ENC_EMB, (t * B + b) * cfg.E + e, 0.5, dtype) for t in range(tmax)],
        "dec_emb": [rng.keep_scale(seed, STREAM_AE_DEC_EMB, (t * B + b) * cfg.E + e, 0.5, dtype) for t in range(tmax + 1)],
        "out": [rng.keep_scale(seed, STREAM_AE_OUT, (t * B + b) * cfg.H + h, cfg.p, dtype) for t in range(tmax + 1)],
    }


def _log_softmax(x):
    m = x.max(axis=1, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True))


def loss_and_grads(cfg, enc_w, dec_w, lut_w, seq, seed=None, dtype=np.float32, masks=None, grad_clip=0.1,
                   weight_decay=1e-6, literal_lookup_grad=False, want_grads=True, keep_logprobs=False):
    """lossFun: returns loss, [g_enc, g_dec, g_lut] (clamped, + weight_decay * params), context."""
    assert cfg.L == 1, "the oracle restates the reference default num_layers = 1"
    enc = A.split_flat(enc_w.astype(dtype), cfg.enc_layout())
    dec = A.split_flat(dec_w.astype(dtype), cfg.dec_layout())
    lut = A.split_flat(lut_w.astype(dtype), cfg.lut_layout())["lookup"]
    seq = np.asarray(seq)
    B = seq.shape[0]
    etok, dtok, tmax = step_tokens(cfg, seq)
    if masks is None and seed is not None:
        masks = build_masks(cfg, seed, B, tmax, dtype)
    H = cfg.H

    def embed(tok, mask):
        pre = lut[tok - 1]
        if mask is not None:
            pre = pre * mask
        return np.tanh(pre)

    # ---- encoder ----
    state = np.zeros((B, cfg.S), dtype=dtype)
    enc_cache, enc_y = [], []
    for t in range(tmax):
        y = embed(etok[t], None if masks is None else masks["enc_emb"][t])
        enc_y.append(y)
        state, cache = A.lstm_cell_forward(cfg, enc, state, y, None)
        enc_cache.append(cache)
    enc_final = state
    # ---- decoder ----
    tg, n = lm_targets(cfg, seq)
    dec_cache, dec_y, hd, logps = [], [], [], []
    loss = 0.0
    for t in range(tmax + 1):
        y = embed(dtok[t], None if masks is None else masks["dec_emb"][t])
        dec_y.append(y)
        state, cache = A.lstm_cell_forward(cfg, dec, state, y, None)
        dec_cache.append(cache)
        top = state[:, (2 * cfg.L - 1) * H:2 * cfg.L * H]
        if masks is not None:
            top = top * masks["out"][t]
        hd.append(top)
        lp = _log_softmax(top @ dec["Wd"].T + dec["bd"])
        logps.append(lp)
        sel = tg[t] != 0
        loss -= float(lp[sel, tg[t][sel] - 1].astype(np.float64).sum())
    loss /= n
    ctx = dict(tmax=tmax, n=n, targets=tg, enc_final=enc_final, dec_final=state)
    if keep_logprobs:
        ctx["logprobs"] = logps
    if not want_grads:
        return loss, None, ctx
    # ---- backward ----
    g_enc = np.zeros(cfg.n_enc, dtype=dtype)
    g_dec = np.zeros(cfg.n_dec, dtype=dtype)
    g_lut = np.zeros(cfg.n_lut, dtype=dtype)
    enc_g = A.split_flat(g_enc, cfg.enc_layout())
    dec_g = A.split_flat(g_dec, cfg.dec_layout())
    lut_g = A.split_flat(g_lut, cfg.lut_layout())["lookup"]

    def embed_backward(tok, y, dy, mask):
        dpre = dy * (1.0 - y * y)
        if mask is not None:
            dpre = dpre * mask
        np.add.at(lut_g, tok - 1, dpre.astype(dtype))

    dstate = np.zeros((B, cfg.S), dtype=dtype)
    for t in reversed(range(tmax + 1)):
        lp = logps[t]
        sel = tg[t] != 0
        dlogits = np.zeros_like(lp)
        # LogSoftMax backward of dlogp = -1/n at the target: (softmax - onehot) / n on predicting rows, 0 elsewhere
        dlogits[sel] = np.exp(lp[sel]) / dtype(n)
        dlogits[sel, tg[t][sel] - 1] -= dtype(1.0) / dtype(n)
        dec_g["Wd"] += dlogits.T @ hd[t]
        dec_g["bd"] += dlogits.sum(axis=0)
        dtop = dlogits @ dec["Wd"]
        if masks is not None:
            dtop = dtop * masks["out"][t]
        dstate = dstate.copy()
        dstate[:, (2 * cfg.L - 1) * H:2 * cfg.L * H] += dtop
        dstate, dx = A.lstm_cell_backward(cfg, dec, dec_g, dec_cache[t], dstate, None)
        embed_backward(dtok[t], dec_y[t], dx, None if masks is None else masks["dec_emb"][t])
    for t in reversed(range(tmax)):
        dstate, dx = A.lstm_cell_backward(cfg, enc, enc_g, enc_cache[t], dstate, None)
        embed_backward(etok[t], enc_y[t], dx, None if masks is None else masks["enc_emb"][t])
    ctx["lut_grad_raw"] = g_lut.copy()
    if literal_lookup_grad:
        g_lut[:] = 0
    grads = [g_enc, g_dec, g_lut]
    if grad_clip is not None:
        grads = [np.clip(g, -grad_clip, grad_clip) for g in grads]
    if weight_decay:
        grads = [g + dtype(weight_decay) * w.astype(dtype) for g, w in zip(grads, (enc_w, dec_w, lut_w))]
    return loss, grads, ctx


def adam_update(x, dx, state, lr, beta1=0.8, beta2=0.999, eps=1e-8):
    """adam() of misc/optim_updates.lua:78-111 (in place on x; state = dict with m, v, t)."""
    dt = x.dtype.type
    if "m" not in state:
        state["t"] = 0
        state["m"] = np.zeros_like(x)
        state["v"] = np.zeros_like(x)
    state["m"] *= dt(beta1)
    state["m"] += dt(1 - beta1) * dx
    state["v"] *= dt(beta2)
    state["v"] += dt(1 - beta2) * dx * dx
    tmp = np.sqrt(state["v"]) + dt(eps)
    state["t"] += 1
    bc1 = 1 - beta1 ** state["t"]
    bc2 = 1 - beta2 ** state["t"]
    step = lr * np.sqrt(bc2) / bc1
    x -= dt(step) * (state["m"] / tmp)


def train_step(cfg, enc_w, dec_w, lut_w, states, seq, lr=1e-5, seed=None, dtype=np.float32, **kw):
    """one iteration of the main loop (001_train_arch1_text_autoencoder.lua:262-341, optim 'adam' :334)."""
    loss, grads, ctx = loss_and_grads(cfg, enc_w, dec_w, lut_w, seq, seed, dtype, **kw)
    for w, g, st in zip((enc_w, dec_w, lut_w), grads, states):
        adam_update(w, g.astype(w.dtype), st, lr)
    return loss, ctx
