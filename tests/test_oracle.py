"""CPU tests that pin the oracle (no GPU): self-consistency, independent PyTorch-CPU cross-checks of the
restated Torch7 semantics, the committed golden vectors, and the host-side integer ABI."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_close, rel_err
from oracle import arch1 as A
from oracle import rng

GOLD = os.path.join(os.path.dirname(__file__), "golden", "arch1_small.npz")


def small():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    return make_golden


def test_golden_vectors_reproduce():
    """The committed fixtures are what the oracle computes today (guards silent oracle drift)."""
    g = np.load(GOLD)
    new = small().compute()
    for k in g.files:
        if g[k].dtype.kind in "iu":
            assert np.array_equal(g[k], new[k]), k
        else:
            np.testing.assert_allclose(new[k], g[k], rtol=1e-5, atol=1e-7, err_msg=k)


def test_fp32_oracle_tracks_fp64_truth():
    g = np.load(GOLD)
    for tag in ("eval", "train"):
        for k in ("loss", "scores", "genc", "gemb", "gmm", "state"):
            assert_close(g[f"{tag}_f32_{k}"], g[f"{tag}_f64_{k}"], 2e-5, f"{tag} {k}")


def test_finite_difference_gradients_fp64():
    mg = small()
    cfg, q, lengths, enc, emb, mm, fc7, labels = mg.make_inputs()
    enc, emb, mm = enc.astype(np.float64), emb.astype(np.float64), mm.astype(np.float64)
    q_ra = A.right_align(q, lengths)
    fv = A.l2_normalize_rows(fc7.astype(np.float64))
    r = np.random.default_rng(0)
    for seed in (None, 5):
        f, g, _, _ = A.jdj(cfg, enc, emb, mm, q_ra, lengths, fv, labels, seed=seed, dtype=np.float64, clamp=None)
        for w, gw in ((enc, g[0]), (emb, g[1]), (mm, g[2])):
            big = np.argsort(-np.abs(gw))[:40]
            for k in r.choice(big, 6, replace=False):
                o, h = w[k], 1e-6
                w[k] = o + h
                fp = A.jdj(cfg, enc, emb, mm, q_ra, lengths, fv, labels, seed=seed, dtype=np.float64)[0]
                w[k] = o - h
                fm = A.jdj(cfg, enc, emb, mm, q_ra, lengths, fv, labels, seed=seed, dtype=np.float64)[0]
                w[k] = o
                fd = (fp - fm) / (2 * h)
                assert abs(fd - gw[k]) <= 1e-5 * max(1e-3, abs(fd)), (seed, k, fd, gw[k])


def test_onehot_linear_equals_gather():
    """nn.Linear(V,E) on a dense one-hot (002_train_baseline.lua:141-144) == column gather + bias."""
    cfg = A.Arch1Config(V=23, E=8, H=4, L=1, I=4, C=4, O=3, T=3)
    r = np.random.default_rng(1)
    emb = A.split_flat(r.standard_normal(cfg.n_emb), cfg.emb_layout())
    words = r.integers(1, cfg.V + 1, 17)
    onehot = np.zeros((17, cfg.V))
    onehot[np.arange(17), words - 1] = 1.0                       # misc/RNNUtils.lua:42-47
    dense = np.tanh(onehot @ emb["We"].T + emb["be"])
    np.testing.assert_allclose(A.embedding_forward(emb, words, None), dense, rtol=1e-12)


def test_cell_matches_torch_lstmcell():
    """Torch7 gate order (i,f,o,g) vs PyTorch's (i,f,g,o): same cell after permuting gate rows."""
    torch = pytest.importorskip("torch")
    cfg = A.Arch1Config(V=5, E=6, H=5, L=2, I=4, C=4, O=3, T=3)
    r = np.random.default_rng(2)
    enc = A.split_flat(r.uniform(-.5, .5, cfg.n_enc), cfg.enc_layout())
    state = r.standard_normal((4, cfg.S))
    x = r.standard_normal((4, cfg.E))
    out, _ = A.lstm_cell_forward(cfg, enc, state, x, None)
    H = cfg.H
    perm = np.concatenate([np.arange(0, 2 * H), np.arange(3 * H, 4 * H), np.arange(2 * H, 3 * H)])
    inp = torch.tensor(x)
    for l in range(cfg.L):
        cell = torch.nn.LSTMCell(cfg.E if l == 0 else H, H).double()
        with torch.no_grad():
            cell.weight_ih.copy_(torch.tensor(enc[f"Wi{l}"][perm]))
            cell.weight_hh.copy_(torch.tensor(enc[f"Wh{l}"][perm]))
            cell.bias_ih.copy_(torch.tensor(enc[f"bi{l}"][perm]))
            cell.bias_hh.copy_(torch.tensor(enc[f"bh{l}"][perm]))
            c0 = torch.tensor(state[:, 2 * l * H:(2 * l + 1) * H])
            h0 = torch.tensor(state[:, (2 * l + 1) * H:(2 * l + 2) * H])
            h1, c1 = cell(inp, (h0, c0))
        np.testing.assert_allclose(out[:, 2 * l * H:(2 * l + 1) * H], c1.numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out[:, (2 * l + 1) * H:(2 * l + 2) * H], h1.numpy(), rtol=1e-10, atol=1e-12)
        inp = h1


def test_criterion_and_rmsprop_match_torch():
    torch = pytest.importorskip("torch")
    r = np.random.default_rng(3)
    s = r.standard_normal((9, 13))
    y = r.integers(1, 14, 9)
    f, d = A.cross_entropy(s, y)
    st = torch.tensor(s, requires_grad=True)
    ft = torch.nn.functional.cross_entropy(st, torch.tensor(y - 1))
    ft.backward()
    assert abs(f - ft.item()) < 1e-12
    np.testing.assert_allclose(d, st.grad.numpy(), rtol=1e-10, atol=1e-14)
    # optim.rmsprop == torch.optim.RMSprop(alpha=.99, eps=1e-8, centered=False): eps outside the sqrt
    x = r.standard_normal(50)
    xt = torch.tensor(x.copy(), requires_grad=True)
    opt = torch.optim.RMSprop([xt], lr=3e-4, alpha=0.99, eps=1e-8)
    m = np.zeros_like(x)
    for _ in range(3):
        g = r.standard_normal(50)
        A.rmsprop_update(x, g, m, 3e-4)
        xt.grad = torch.tensor(g)
        opt.step()
    np.testing.assert_allclose(x, xt.detach().numpy(), rtol=1e-12)


def test_packed_recurrence_equals_padded_masked():
    """The CUDA path's formulation (padded rows, activity mask, no sort) is the same function as the
    reference's sorted/packed one: check on the oracle's own primitives in fp64."""
    mg = small()
    cfg, q, lengths, enc, emb, mm, fc7, labels = mg.make_inputs()
    q_ra = A.right_align(q, lengths)
    encd = A.split_flat(enc.astype(np.float64), cfg.enc_layout())
    embd = A.split_flat(emb.astype(np.float64), cfg.emb_layout())
    _, ctx = A.forward(cfg, enc, emb, mm, q_ra, lengths, A.l2_normalize_rows(fc7.astype(np.float64)), None, np.float64)
    B = q_ra.shape[0]
    state = np.zeros((B, cfg.S))
    for t in range(cfg.T):
        act = (t >= cfg.T - lengths)
        x = np.zeros((B, cfg.E))
        x[act] = A.embedding_forward(embd, q_ra[act, t], None)
        new, _ = A.lstm_cell_forward(cfg, encd, state, x, None)
        state = np.where(act[:, None], new, 0.0)
    np.testing.assert_allclose(state, ctx["tv_q"], rtol=1e-12, atol=1e-14)


def test_pack_batch_c_abi_bit_exact():
    """nvqa_right_align / nvqa_pack_batch (host C, no GPU needed) == oracle, including ties and ragged lengths."""
    from novel_vqa_b200 import right_align, pack_batch
    r = np.random.default_rng(4)
    for B, T in ((1, 1), (6, 7), (37, 26), (500, 26)):
        lengths = r.integers(1, T + 1, B).astype(np.int32)
        if B == 500:
            lengths[:] = T                                   # the BASELINE configuration: all full length
        q = np.zeros((B, T), dtype=np.int32)
        for b in range(B):
            q[b, :lengths[b]] = r.integers(1, 1000, lengths[b])
        ra = right_align(q, lengths)
        assert np.array_equal(ra, A.right_align(q, lengths))
        words, sizes, sidx, inv = pack_batch(ra, lengths)
        ow, os_, osi, oinv = A.sort_encoding_right_align(ra, lengths)
        assert np.array_equal(words, ow) and np.array_equal(sizes, os_)
        assert np.array_equal(sidx, osi + 1) and np.array_equal(inv, oinv + 1)
        assert int(sizes.sum()) == int(lengths.sum()) and np.all(np.diff(sizes) >= 0)
    g = np.load(GOLD)
    words, sizes, sidx, inv = pack_batch(g["q_ra"], g["lengths"])
    assert np.array_equal(words, g["words"]) and np.array_equal(sizes, g["batch_sizes"])
    assert np.array_equal(sidx, g["sort_index"]) and np.array_equal(inv, g["sort_index_inverse"])


def test_rng_statistics_and_determinism():
    idx = np.arange(1 << 16)
    m = rng.keep_scale(123, rng.STREAM_EMB, idx, 0.5)
    assert set(np.unique(m)) == {0.0, 2.0}
    assert abs(m.mean() - 1.0) < 0.02
    assert np.array_equal(m, rng.keep_scale(123, rng.STREAM_EMB, idx, 0.5))
    assert not np.array_equal(m, rng.keep_scale(124, rng.STREAM_EMB, idx, 0.5))
    assert not np.array_equal(m, rng.keep_scale(123, rng.STREAM_HEAD, idx, 0.5))


def test_multiple_choice_select():
    s = np.array([0.1, 0.9, 0.3, 0.9])
    assert A.multiple_choice_select(s, [3, 0, 4, 2, 0]) == 4      # first max among candidates in list order
    assert A.argmax_first(np.array([[1.0, 3.0, 3.0]]))[0] == 2


def test_multiple_choice_select_c_abi_bit_exact():
    """nvqa_mc_select (host, 004_eval_model.lua:257-271) against the oracle, incl. ties and padded candidate lists."""
    import novel_vqa_b200 as nv
    r = np.random.default_rng(3)
    n, O, K = 300, 1000, 18
    scores = r.standard_normal((n, O)).astype(np.float32)
    scores[:, ::7] = 0.25                                      # plenty of exact ties
    mc = np.zeros((n, K), dtype=np.int32)
    for i in range(n):
        k = r.integers(1, K + 1)
        mc[i, :k] = r.choice(np.arange(1, O + 1), k, replace=False)
    got = nv.mc_select(scores, mc)
    want = [A.multiple_choice_select(scores[i], mc[i]) for i in range(n)]
    assert got.tolist() == want


def test_trainer_variants_finite_differences_fp64():
    """AskipB fusion (misc/netdef.lua:16-25) and the lr_scale / two-block-norm variants of 003_train_ae_based_{wp,ef}.lua."""
    cfg = A.Arch1Config(V=40, E=8, H=8, L=1, I=12, C=8, O=6, T=5, fusion_skip=True)
    r = np.random.RandomState(0)
    enc, emb, mm = (r.uniform(-.4, .4, n) for n in (cfg.n_enc, cfg.n_emb, cfg.n_mm))
    B = 5
    lens = np.array([5, 3, 1, 4, 2])
    q = np.zeros((B, cfg.T), dtype=np.int64)
    for b in range(B):
        q[b, cfg.T - lens[b]:] = r.randint(1, cfg.V + 1, lens[b])
    fv = A.l2_normalize_rows(np.abs(r.randn(B, cfg.I)), split=4)
    assert np.allclose((fv[:, :4] ** 2).sum(1), 1) and np.allclose((fv[:, 4:] ** 2).sum(1), 1)
    lab = r.randint(1, cfg.O + 1, B)
    f, g, _, _ = A.jdj(cfg, enc, emb, mm, q, lens, fv, lab, seed=3, dtype=np.float64, clamp=None)
    eps = 1e-6
    for w, gw in ((enc, g[0]), (mm, g[2])):
        for i in r.choice(len(w), 8, replace=False):
            w[i] += eps
            fp = A.jdj(cfg, enc, emb, mm, q, lens, fv, lab, seed=3, dtype=np.float64, clamp=None)[0]
            w[i] -= 2 * eps
            fm = A.jdj(cfg, enc, emb, mm, q, lens, fv, lab, seed=3, dtype=np.float64, clamp=None)[0]
            w[i] += eps
            assert abs((fp - fm) / (2 * eps) - gw[i]) <= 1e-6 + 1e-4 * abs(gw[i])
    # lr_scale multiplies the encoder and embedding blocks before the clamp, not the multimodal block
    _, gs, _, _ = A.jdj(cfg, enc, emb, mm, q, lens, fv, lab, seed=3, dtype=np.float64, clamp=None, lr_scale=0.1)
    assert np.allclose(gs[0], 0.1 * g[0]) and np.allclose(gs[1], 0.1 * g[1]) and np.allclose(gs[2], g[2], rtol=1e-12, atol=0)


def test_rng_expectation_is_one_for_any_p():
    """the 8-bit keep threshold quantises the keep probability to k/256; the multiplier is its exact reciprocal,
    so E[mask] = 1 also for p that is not a multiple of 1/256 (0.3 = the reference's -drop_prob_ae option range), and a p
    above 255/256 still keeps 1/256 of the elements instead of dropping everything."""
    idx = np.arange(1 << 18)
    for p in (0.3, 0.1, 0.75, 0.999):
        m = rng.keep_scale(7, rng.STREAM_EMB, idx, p)
        kept = m[m > 0]
        k = 256 - min(255, int(np.float32(p) * np.float32(256) + np.float32(0.5)))
        assert kept.size and np.all(kept == np.float32(256.0 / k))
        assert abs((m > 0).mean() - k / 256) < 0.01
        assert abs(m.mean() - 1.0) < (0.03 if p < 0.9 else 0.25)


@pytest.mark.parametrize("variant", ["faithful", "gather"])
@pytest.mark.parametrize("seed", [None, 5])
def test_torch_cpu_port_matches_oracle(variant, seed):
    """oracle/torch_cpu.py (the timed CPU baseline of bench.py: dense one-hot Linear, un-shared clones, autograd) computes
    the same loss, scores, clamped gradients and RMSprop update as oracle/arch1.py -- ragged lengths, evaluate and training
    mode (shared hash masks)."""
    from oracle import torch_cpu
    mg = small()
    cfg = A.Arch1Config(**mg.CFG)
    g = np.load(GOLD)
    q, ln, lab = g["q_ra"], g["lengths"], g["labels"]
    fv = A.l2_normalize_rows(g["fc7"])
    port = torch_cpu.TorchCpuArch1(cfg, g["enc"], g["emb"], g["mm"], variant)
    f, grads, scores = port.jdj(q, ln, fv, lab, seed=seed)
    f_ref, g_ref, s_ref, _ = A.jdj(cfg, g["enc"], g["emb"], g["mm"], q, ln, fv, lab, seed=seed)
    assert abs(f - f_ref) <= 1e-5 * abs(f_ref)
    assert_close(scores.numpy(), s_ref, 1e-5, "scores")
    assert_close(grads.numpy(), np.concatenate(g_ref), 1e-5, "clamped gradients [encoder | embedding | multimodal]")
    # one optimizer step (the port's 6-pass RMSprop over the joined vector == rmsprop_update per block)
    port2 = torch_cpu.TorchCpuArch1(cfg, g["enc"], g["emb"], g["mm"], variant)
    port2.step((q, ln, fv, lab), 3e-4, seed=seed)
    w = [g["enc"].copy(), g["emb"].copy(), g["mm"].copy()]
    A.train_step(cfg, w[0], w[1], w[2], [np.zeros_like(x) for x in w], (q, ln, fv, lab), 3e-4, seed=seed)
    ref_x, x0 = np.concatenate(w), np.concatenate([g["enc"], g["emb"], g["mm"]])
    big = np.abs(np.concatenate(g_ref)) > 1e-5           # lr * g / (0.1 |g| + eps) is ill-conditioned where |g| ~ eps
    assert_close((port2.x.numpy() - x0)[big], (ref_x - x0)[big], 5e-3, "RMSprop update")
