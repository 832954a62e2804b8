"""CPU tests that pin the text-autoencoder oracle (oracle/ae.py) without Torch7: finite differences in fp64, an
independent PyTorch-autograd restatement, the criterion's target rule on hand-made cases, Adam against torch.optim's
building blocks, and the committed golden vectors."""
import os

import numpy as np
import torch

from conftest import assert_close
from oracle import ae as AE

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ae_small.npz")


def small(seed=0, B=6, T=5, V=30, E=8, H=8):
    cfg = AE.AEConfig(V=V, E=E, H=H, L=1, T=T)
    r = np.random.RandomState(seed)
    enc = r.uniform(-.3, .3, cfg.n_enc)
    dec = r.uniform(-.3, .3, cfg.n_dec)
    lut = r.uniform(-.5, .5, cfg.n_lut)
    lens = r.randint(1, T + 1, B)
    lens[0] = T
    seq = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        seq[b, :lens[b]] = r.randint(1, V + 1, lens[b])
    return cfg, enc, dec, lut, seq


def test_lm_targets_rule():
    cfg = AE.AEConfig(V=9, T=4)
    seq = np.array([[3, 4, 0, 0], [1, 2, 3, 4], [5, 0, 0, 0]])
    tg, n = AE.lm_targets(cfg, seq)
    # first null -> END (= V+1 = 10), later nulls skipped; a full-length row predicts END at step T+1
    assert tg.T.tolist() == [[3, 4, 10, 0, 0], [1, 2, 3, 4, 10], [5, 10, 0, 0, 0]]
    assert n == 3 + 5 + 2
    enc, dec, tmax = AE.step_tokens(cfg, seq)
    assert tmax == 4 and enc[2].tolist() == [1, 3, 1]            # nulls are fed as token 1, unmasked
    assert dec[0].tolist() == [10, 10, 10] and dec[1].tolist() == [3, 1, 5]


def test_finite_difference_gradients_fp64():
    cfg, enc, dec, lut, seq = small()
    f, g, _ = AE.loss_and_grads(cfg, enc, dec, lut, seq, seed=7, dtype=np.float64, grad_clip=None, weight_decay=0)
    r = np.random.RandomState(1)
    eps = 1e-6
    for w, gw in ((enc, g[0]), (dec, g[1]), (lut, g[2])):
        for i in r.choice(len(w), 10, replace=False):
            w[i] += eps
            fp, _, _ = AE.loss_and_grads(cfg, enc, dec, lut, seq, seed=7, dtype=np.float64, want_grads=False)
            w[i] -= 2 * eps
            fm, _, _ = AE.loss_and_grads(cfg, enc, dec, lut, seq, seed=7, dtype=np.float64, want_grads=False)
            w[i] += eps
            assert abs((fp - fm) / (2 * eps) - gw[i]) <= 1e-6 + 1e-4 * abs(gw[i])


def test_matches_independent_torch_autograd():
    """Same model written with torch ops + autograd (torch's own LSTM cell with the gate rows permuted)."""
    cfg, enc, dec, lut, seq = small(seed=3, B=5, T=6)
    masks = AE.build_masks(cfg, 11, seq.shape[0], int((seq != 0).sum(1).max()), np.float64)
    f, g, ctx = AE.loss_and_grads(cfg, enc, dec, lut, seq, dtype=np.float64, masks=masks, grad_clip=None, weight_decay=0)
    H, V1 = cfg.H, cfg.V + 1
    te, td, tl = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (enc, dec, lut))

    def core(flat):
        o, out = 0, {}
        for name, shape in cfg.core_layout():
            n = int(np.prod(shape))
            out[name] = flat[o:o + n].view(*shape)
            o += n
        return out, o

    pe, _ = core(te)
    pd, o = core(td)
    Wd = td[o:o + V1 * H].view(V1, H)
    bd = td[o + V1 * H:]
    table = tl.view(V1, cfg.E)
    # torch.nn.LSTMCell orders its gate rows i,f,g,o; the reference orders them i,f,o,g (LSTM_encoder.lua:36-43)
    perm = torch.cat([torch.arange(0, 2 * H), torch.arange(3 * H, 4 * H), torch.arange(2 * H, 3 * H)])

    def cell(p, x, h, c):
        return torch._VF.lstm_cell(x, (h, c), p["Wi0"][perm], p["Wh0"][perm], p["bi0"][perm], p["bh0"][perm])

    B = seq.shape[0]
    tmax = ctx["tmax"]
    h = torch.zeros(B, H, dtype=torch.float64)
    c = torch.zeros(B, H, dtype=torch.float64)
    tok = torch.tensor(np.where(seq == 0, 1, seq)) - 1
    for t in range(tmax):
        x = torch.tanh(table[tok[:, t]] * torch.tensor(masks["enc_emb"][t]))
        h, c = cell(pe, x, h, c)
    tg = torch.tensor(ctx["targets"])
    loss = 0.0
    for t in range(tmax + 1):
        idx = torch.full((B,), V1 - 1, dtype=torch.long) if t == 0 else tok[:, t - 1]
        x = torch.tanh(table[idx] * torch.tensor(masks["dec_emb"][t]))
        h, c = cell(pd, x, h, c)
        lp = torch.log_softmax((h * torch.tensor(masks["out"][t])) @ Wd.T + bd, dim=1)
        sel = tg[t] != 0
        loss = loss - lp[sel, tg[t][sel] - 1].sum()
    loss = loss / ctx["n"]
    loss.backward()
    assert abs(loss.item() - f) <= 1e-12 * max(1, abs(f))
    for a, b, what in ((te.grad, g[0], "encoder"), (td.grad, g[1], "decoder"), (tl.grad, g[2], "lookup")):
        assert_close(b, a.numpy(), 1e-10, what)


def test_fp32_tracks_fp64_and_literal_lookup_mode():
    cfg, enc, dec, lut, seq = small(seed=5)
    f64, g64, _ = AE.loss_and_grads(cfg, enc, dec, lut, seq, seed=3, dtype=np.float64)
    f32, g32, c32 = AE.loss_and_grads(cfg, enc.astype(np.float32), dec.astype(np.float32), lut.astype(np.float32), seq, seed=3)
    assert abs(f32 - f64) <= 1e-5 * abs(f64)
    for a, b in zip(g32, g64):
        assert_close(a, b, 1e-4, "fp32 vs fp64 gradients")
    # literal reference behaviour: the LookupTable block of grad_params only carries the weight-decay term
    _, gl, _ = AE.loss_and_grads(cfg, enc, dec, lut, seq, seed=3, dtype=np.float64, literal_lookup_grad=True)
    assert np.allclose(gl[2], 1e-6 * lut, rtol=0, atol=1e-15)
    assert np.array_equal(gl[0], g64[0]) and np.array_equal(gl[1], g64[1])


def test_adam_matches_torch_building_blocks():
    r = np.random.RandomState(0)
    x = r.randn(1000)
    xt = torch.tensor(x.copy())
    st = {}
    m = torch.zeros(1000, dtype=torch.float64)
    v = torch.zeros(1000, dtype=torch.float64)
    for t in range(1, 6):
        g = r.randn(1000) * 0.1
        AE.adam_update(x, g, st, 1e-3)
        gt = torch.tensor(g)
        m = 0.8 * m + 0.2 * gt
        v = 0.999 * v + 0.001 * gt * gt
        step = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.8 ** t)
        xt = xt - step * m / (v.sqrt() + 1e-8)          # eps OUTSIDE the sqrt, bias corrections in the step size
    assert np.allclose(x, xt.numpy(), rtol=1e-12, atol=1e-14)


def test_golden_vectors_reproduce():
    g = np.load(GOLD)
    cfg = AE.AEConfig(**{k: int(g["cfg_" + k]) for k in ("V", "E", "H", "L", "T")})
    f, grads, ctx = AE.loss_and_grads(cfg, g["enc"], g["dec"], g["lut"], g["seq"], seed=int(g["seed"]), keep_logprobs=True)
    assert abs(f - float(g["loss"])) <= 1e-6 * abs(f)
    for a, name in zip(grads, ("g_enc", "g_dec", "g_lut")):
        assert_close(a, g[name], 1e-5, name)
    assert np.array_equal(ctx["targets"], g["targets"]) and ctx["n"] == int(g["n"])
    assert_close(ctx["logprobs"][1], g["logprobs1"], 1e-5, "logprobs of decoder step 2")
