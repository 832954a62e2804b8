"""CPU tests that pin the arch2 oracle (oracle/arch2.py) without Torch7: finite differences in fp64, an independent
PyTorch-autograd restatement of 003_train_vqa_arch2/002_train_baseline.lua:277-333 + misc/Encoder_lstm.lua:152-263
(torch's own LSTM cell with the gate rows permuted, nn.functional.embedding, cross_entropy), the step-token rule on
hand-made cases, and the RMSprop-with-weight-decay update."""
import numpy as np
import torch

from conftest import assert_close
from oracle import arch1 as A
from oracle import arch2 as A2


def small(seed=0, B=6, T=5, V=20, E=8, H=8, L=1, I=7, O=5):
    cfg = A2.Arch2Config(V=V, E=E, H=H, L=L, I=I, O=O, T=T)
    r = np.random.RandomState(seed)
    cnn = r.uniform(-.3, .3, cfg.n_cnn)
    enc = r.uniform(-.3, .3, cfg.n_enc)
    mm = r.uniform(-.3, .3, cfg.n_mm)
    lens = r.randint(1, T, B)                                # every question shorter than T: tmax < T + 2
    lens[0] = T - 1
    seq = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        seq[b, :lens[b]] = r.randint(1, V + 1, lens[b])
    fv = np.maximum(r.standard_normal((B, I)), 0)
    lab = r.randint(1, O + 1, B)
    return cfg, cnn, enc, mm, seq, fv, lab


def test_step_tokens_rule():
    """Encoder_lstm.lua:170-203: image, START (= V + 1), then the columns of seq with nulls replaced by token 1; the loop
    stops being executed once a column is all null (tmax = longest question + 2)."""
    cfg = A2.Arch2Config(V=9, E=4, H=4, L=1, I=3, O=2, T=4)
    seq = np.array([[3, 5, 0, 0], [7, 0, 0, 0]])
    toks, tmax = A2.step_tokens(cfg, seq)
    assert tmax == 4 and toks[0] is None
    assert toks[1].tolist() == [10, 10]
    assert toks[2].tolist() == [3, 7] and toks[3].tolist() == [5, 1]      # the null of row 1 became token 1
    assert toks[4] is None and toks[5] is None
    full = np.array([[1, 2, 3, 4]])
    assert A2.step_tokens(cfg, full)[1] == 6


def test_finite_difference_gradients_fp64():
    for L, seed in ((1, None), (2, 9)):                       # evaluate mode, and training mode with inter-layer dropout
        cfg, cnn, enc, mm, seq, fv, lab = small(L=L)
        f, g, _, _ = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=seed, dtype=np.float64, clamp=None)
        r = np.random.RandomState(1)
        eps = 1e-6
        for w, gw in ((cnn, g[0]), (enc, g[1]), (mm, g[2])):
            for i in r.choice(len(w), 12, replace=False):
                w[i] += eps
                fp = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=seed, dtype=np.float64, clamp=None)[0]
                w[i] -= 2 * eps
                fm = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=seed, dtype=np.float64, clamp=None)[0]
                w[i] += eps
                assert abs((fp - fm) / (2 * eps) - gw[i]) <= 1e-6 + 1e-4 * abs(gw[i])


def test_matches_independent_torch_autograd():
    cfg, cnn, enc, mm, seq, fv, lab = small(seed=3, B=5, T=6, L=2)
    f, g, scores, ctx = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=None, dtype=np.float64, clamp=None)
    H, E, V1 = cfg.H, cfg.E, cfg.V + 1
    tc, te, tm = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (cnn, enc, mm))
    Wcnn, bcnn = tc[:E * cfg.I].view(E, cfg.I), tc[E * cfg.I:]
    o, layers = 0, []
    for l in range(cfg.L):
        n_in = E if l == 0 else H
        p = []
        for shape in ((4 * H, n_in), (4 * H,), (4 * H, H), (4 * H,)):
            n = int(np.prod(shape))
            p.append(te[o:o + n].view(*shape))
            o += n
        layers.append(p)
    table = te[o:].view(V1, E)
    Wc, bc = tm[:cfg.O * H].view(cfg.O, H), tm[cfg.O * H:]
    # torch orders the gate rows i,f,g,o; the reference i,f,o,g (LSTM_encoder.lua:36-43)
    perm = torch.cat([torch.arange(0, 2 * H), torch.arange(3 * H, 4 * H), torch.arange(2 * H, 3 * H)])
    B = seq.shape[0]
    longest = int((seq != 0).sum(1).max())
    hs = [torch.zeros(B, H, dtype=torch.float64) for _ in range(cfg.L)]
    cs = [torch.zeros(B, H, dtype=torch.float64) for _ in range(cfg.L)]
    tok = torch.tensor(np.where(seq == 0, 1, seq)) - 1
    for t in range(longest + 2):                             # image, START, the words up to the longest question
        if t == 0:
            x = torch.tensor(fv) @ Wcnn.T + bcnn
        elif t == 1:
            x = table[torch.full((B,), V1 - 1, dtype=torch.long)]
        else:
            x = torch.nn.functional.embedding(tok[:, t - 2], table)
        for l, (Wi, bi, Wh, bh) in enumerate(layers):
            hs[l], cs[l] = torch._VF.lstm_cell(x, (hs[l], cs[l]), Wi[perm], Wh[perm], bi[perm], bh[perm])
            x = hs[l]
    assert ctx["tmax"] == longest + 2
    logits = hs[-1] @ Wc.T + bc
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor(lab) - 1)
    loss.backward()
    assert abs(loss.item() - f) <= 1e-12 * max(1, abs(f))
    assert_close(scores, logits.detach().numpy(), 1e-12, "scores")
    for a, b, what in ((tc.grad, g[0], "cnn"), (te.grad, g[1], "encoder"), (tm.grad, g[2], "multimodal")):
        assert_close(b, a.numpy(), 1e-10, what)


def test_train_step_is_rmsprop_with_weight_decay_after_the_clamp():
    """003_train_vqa_arch2/002_train_baseline.lua:197,326-331: g = clamp(dJ); optim.rmsprop adds weightDecay * x to g."""
    cfg, cnn, enc, mm, seq, fv, lab = small(seed=5)
    w = [a.astype(np.float32) for a in (cnn * 40, enc, mm)]  # large cnn weights: some gradients hit the clamp
    f, g, _, _ = A2.jdj(cfg, w[0], w[1], w[2], seq, fv, lab, seed=None)
    ms = [np.zeros_like(a) for a in w]
    w2 = [a.copy() for a in w]
    A2.train_step(cfg, w2[0], w2[1], w2[2], ms, (seq, fv, lab), 3e-4, seed=None)
    for a, b, gr in zip(w, w2, g):
        tp = torch.tensor(a.copy(), requires_grad=True)
        opt = torch.optim.RMSprop([tp], lr=3e-4, alpha=0.99, eps=1e-8, weight_decay=1e-4)
        tp.grad = torch.tensor(np.clip(gr, -10, 10).astype(np.float32))
        opt.step()
        assert_close(b, tp.detach().numpy(), 1e-6, "rmsprop + weight decay")


def test_literal_reference_flags_lookup_gradient_and_stale_h0():
    """DESIGN 2 / SURVEY App. C-5.  literal_lookup_grad: the LookupTable block of the gradient is zero, everything else
    unchanged.  h0_top: the forward starts the top layer from the given tensor; checked against torch autograd (the
    gradient w.r.t. the weights flows through the non-zero initial state, the state itself gets no gradient)."""
    cfg, cnn, enc, mm, seq, fv, lab = small(seed=4, B=5, T=6, L=1)
    f0, g0, s0, ctx0 = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=None, dtype=np.float64, clamp=None)
    f1, g1, s1, _ = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=None, dtype=np.float64, clamp=None, literal_lookup_grad=True)
    n_core = cfg.n_enc - (cfg.V + 1) * cfg.E
    assert f0 == f1 and np.array_equal(s0, s1)
    assert np.array_equal(g1[0], g0[0]) and np.array_equal(g1[2], g0[2]) and np.array_equal(g1[1][:n_core], g0[1][:n_core])
    assert not g1[1][n_core:].any() and g0[1][n_core:].any()
    # stale h0 = the head's gradInput of the previous step
    h0 = ctx0["dz"] * 50                                      # scaled up so that it visibly changes the forward
    f2, g2, s2, _ = A2.jdj(cfg, cnn, enc, mm, seq, fv, lab, seed=None, dtype=np.float64, clamp=None, h0_top=h0)
    assert np.abs(s2 - s0).max() > 1e-6
    H, E, V1 = cfg.H, cfg.E, cfg.V + 1
    tc, te, tm = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (cnn, enc, mm))
    Wcnn, bcnn = tc[:E * cfg.I].view(E, cfg.I), tc[E * cfg.I:]
    o, p = 0, []
    for shape in ((4 * H, E), (4 * H,), (4 * H, H), (4 * H,)):
        n = int(np.prod(shape))
        p.append(te[o:o + n].view(*shape))
        o += n
    table = te[o:].view(V1, E)
    Wc, bc = tm[:cfg.O * H].view(cfg.O, H), tm[cfg.O * H:]
    perm = torch.cat([torch.arange(0, 2 * H), torch.arange(3 * H, 4 * H), torch.arange(2 * H, 3 * H)])
    B = seq.shape[0]
    longest = int((seq != 0).sum(1).max())
    h, c = torch.tensor(h0), torch.zeros(B, H, dtype=torch.float64)
    tok = torch.tensor(np.where(seq == 0, 1, seq)) - 1
    for t in range(longest + 2):
        x = (torch.tensor(fv) @ Wcnn.T + bcnn) if t == 0 else table[torch.full((B,), V1 - 1, dtype=torch.long)] if t == 1 \
            else torch.nn.functional.embedding(tok[:, t - 2], table)
        h, c = torch._VF.lstm_cell(x, (h, c), p[0][perm], p[2][perm], p[1][perm], p[3][perm])
    loss = torch.nn.functional.cross_entropy(h @ Wc.T + bc, torch.tensor(lab) - 1)
    loss.backward()
    assert abs(loss.item() - f2) <= 1e-12 * max(1, abs(f2))
    for a, b, what in ((tc.grad, g2[0], "cnn"), (te.grad, g2[1], "encoder"), (tm.grad, g2[2], "multimodal")):
        assert_close(b, a.numpy(), 1e-10, what + " with a non-zero initial state")
    # train_step(literal=True) threads the carry: step 2 differs from a non-literal step 2, and a batch-size change resets
    w = [a.astype(np.float32) for a in (cnn, enc, mm)]
    wl = [a.copy() for a in w]
    ms, msl, carry = [np.zeros_like(a) for a in w], [np.zeros_like(a) for a in w], {}
    fa = [A2.train_step(cfg, w[0], w[1], w[2], ms, (seq, fv, lab), 3e-4)[0] for _ in range(2)]
    fl = [A2.train_step(cfg, wl[0], wl[1], wl[2], msl, (seq, fv, lab), 3e-4, literal=True, carry=carry)[0] for _ in range(2)]
    assert fa[0] == fl[0] and fa[1] != fl[1]
    assert carry["B"] == B and carry["dz"].shape == (B, H)
    A2.train_step(cfg, wl[0], wl[1], wl[2], msl, (seq[:3], fv[:3], lab[:3]), 3e-4, literal=True, carry=carry)
    assert carry["B"] == 3
