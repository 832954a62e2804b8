"""GPU parity tests of the text-autoencoder path (BASELINE config 5), through the C ABI, against
oracle/ae.py and the committed golden vectors.  Integer outputs (targets via the loss count) exact; fp32 modes within
1e-4 relative (rel-L2 and rel-max per tensor), the bf16-operand mode within 1e-2."""
import os

import numpy as np
import pytest

from conftest import assert_close, rel_err
from oracle import ae as AE

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ae_small.npz")
PRECISIONS = [("fp32_simt", 0, 1e-4), ("bf16x2", 3, 1e-4), ("bf16x3", 1, 1e-4), ("bf16", 2, 1e-2)]


def nv():
    import novel_vqa_b200 as nv_
    if nv_.device_count() == 0:
        pytest.fail("no sm_100 device visible: GPU tests must run on the B200 box (no CPU fallback)")
    return nv_


def make(nvm, cfg, enc, dec, lut, prec):
    m = nvm.AEModel(cfg, precision=prec)
    m.set_params(nvm.BLOCK_AE_ENCODER, enc)
    m.set_params(nvm.BLOCK_AE_DECODER, dec)
    m.set_params(nvm.BLOCK_AE_LOOKUP, lut)
    return m


def raw_grads(nvm, m):
    return [m.get_grads(b) for b in (nvm.BLOCK_AE_ENCODER, nvm.BLOCK_AE_DECODER, nvm.BLOCK_AE_LOOKUP)]


@pytest.mark.parametrize("name,prec,tol", PRECISIONS)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_small_step_against_golden_and_oracle(name, prec, tol, mode):
    nvm = nv()
    g = np.load(GOLD)
    ocfg = AE.AEConfig(**{k: int(g["cfg_" + k]) for k in ("V", "E", "H", "L", "T")})
    seq, B = g["seq"], g["seq"].shape[0]
    cfg = nvm.AEConfig(V=ocfg.V, E=ocfg.E, H=ocfg.H, L=ocfg.L, T=ocfg.T, B=B)
    m = make(nvm, cfg, g["enc"], g["dec"], g["lut"], prec)
    for blk, key in ((nvm.BLOCK_AE_ENCODER, "enc"), (nvm.BLOCK_AE_DECODER, "dec"), (nvm.BLOCK_AE_LOOKUP, "lut")):
        assert np.array_equal(m.get_params(blk), g[key])                    # checkpoint layout round trip
    seed = int(g["seed"])
    if mode == "train":
        f, grads, ctx = float(g["loss"]), [g["g_enc"], g["g_dec"], g["g_lut"]], None
        lp1, encf = g["logprobs1"], g["enc_final"]
        raw = AE.loss_and_grads(ocfg, g["enc"], g["dec"], g["lut"], seq, seed=seed, grad_clip=None, weight_decay=0)[1]
    else:
        f, _, ctx = AE.loss_and_grads(ocfg, g["enc"], g["dec"], g["lut"], seq, seed=None, keep_logprobs=True,
                                      grad_clip=None, weight_decay=0)
        raw = AE.loss_and_grads(ocfg, g["enc"], g["dec"], g["lut"], seq, seed=None, grad_clip=None, weight_decay=0)[1]
        lp1, encf = ctx["logprobs"][1], ctx["enc_final"]
    m.set_batch_host(seq, g["lengths"])
    m.forward(nvm.MODE_TRAIN if mode == "train" else nvm.MODE_EVAL, seed)
    assert abs(m.loss() - f) <= tol * abs(f)
    assert_close(m.logprobs(1, B), lp1, tol, "log-probs of decoder step 2")
    assert_close(m.state(B), encf, tol, "encoder output state [c|h]")
    m.backward()
    for got, want, what in zip(raw_grads(nvm, m), raw, ("encoder", "decoder", "lookup")):
        assert_close(got, want, tol, f"{what} gradient (raw)")
    with pytest.raises(nvm.NvqaError):
        m.backward()                                                          # log-probs were consumed in place
    m.close()


@pytest.mark.parametrize("name,prec,tol", [("fp32_simt", 0, 1e-4), ("bf16x2", 3, 1e-4)])
def test_training_trajectory_against_golden(name, prec, tol):
    """three iterations of lossFun + clamp + weight decay + adam (lr 1e-3 so that the update is visible)."""
    nvm = nv()
    g = np.load(GOLD)
    ocfg = AE.AEConfig(**{k: int(g["cfg_" + k]) for k in ("V", "E", "H", "L", "T")})
    B = g["seq"].shape[0]
    m = make(nvm, nvm.AEConfig(V=ocfg.V, E=ocfg.E, H=ocfg.H, L=ocfg.L, T=ocfg.T, B=B), g["enc"], g["dec"], g["lut"], prec)
    losses = []
    for i in range(3):
        m.set_batch_host(g["seq"], g["lengths"])
        m.forward(nvm.MODE_TRAIN, int(g["seed"]) + i)
        m.backward()
        m.adam_step(lr=1e-3)
        losses.append(m.loss())
    assert np.allclose(losses, g["traj_losses"], rtol=tol, atol=0)
    for blk, key in ((nvm.BLOCK_AE_ENCODER, "traj_enc"), (nvm.BLOCK_AE_DECODER, "traj_dec"), (nvm.BLOCK_AE_LOOKUP, "traj_lut")):
        # compare the UPDATE (params moved by ~lr per step), not just the parameters
        key0 = key.replace("traj_", "")
        assert_close(m.get_params(blk) - g[key0], g[key] - g[key0], 20 * tol, f"{key} update after 3 adam steps")
    m.close()


@pytest.mark.parametrize("name,prec,tol", [("fp32_simt", 0, 1e-4), ("bf16x2", 3, 1e-4), ("bf16", 2, 1e-2)])
def test_medium_ragged_batch_live_oracle(name, prec, tol):
    """B not a multiple of the 128-row tiles, vocabulary not a multiple of 4, all-short batch (tmax < T)."""
    nvm = nv()
    cfg = nvm.AEConfig(V=1002, E=64, H=128, L=1, T=9, B=200)
    ocfg = AE.AEConfig(V=cfg.V, E=cfg.E, H=cfg.H, L=1, T=cfg.T)
    enc, dec, lut = nvm.synth_params_ae(cfg, seed=5)
    seq, lens = nvm.synth_batch_ae(cfg, 173, seed=6, min_len=1)
    seq[:, 7:] = 0                                                          # tmax = 7 < T
    lens = np.minimum(lens, 7).astype(np.int32)
    f, grads, ctx = AE.loss_and_grads(ocfg, enc, dec, lut, seq, seed=99, grad_clip=None, weight_decay=0)
    assert ctx["tmax"] == 7
    m = make(nvm, cfg, enc, dec, lut, prec)
    m.set_batch_host(seq, lens)
    m.forward(nvm.MODE_TRAIN, 99)
    assert abs(m.loss() - f) <= tol * abs(f)
    m.backward()
    for got, want, what in zip(raw_grads(nvm, m), grads, ("encoder", "decoder", "lookup")):
        assert_close(got, want, tol, f"{what} gradient")
    m.close()


def test_full_size_config5_step():
    """BASELINE config 5: B = 1000, T = 16, V = 20000 (+1), E = H = 512; default fp32-parity mode (bf16x2) against the
    fp32 oracle, plus size-independent properties of the criterion."""
    nvm = nv()
    cfg = nvm.AEConfig()
    ocfg = AE.AEConfig()
    enc, dec, lut = nvm.synth_params_ae(cfg, seed=123)
    seq, lens = nvm.synth_batch_ae(cfg, cfg.B, seed=123, min_len=4)
    f, grads, ctx = AE.loss_and_grads(ocfg, enc, dec, lut, seq, seed=7, grad_clip=None, weight_decay=0)
    m = make(nvm, cfg, enc, dec, lut, nvm.PREC_BF16X2)
    m.set_batch_host(seq, lens)
    m.forward(nvm.MODE_TRAIN, 7)
    loss = m.loss()
    assert abs(loss - f) <= 1e-4 * abs(f)
    assert abs(loss - np.log(cfg.V + 1)) < 0.05                              # near-uniform predictions at random init
    lp = m.logprobs(3, cfg.B)
    assert np.allclose(np.exp(lp.astype(np.float64)).sum(axis=1), 1.0, atol=1e-4)   # rows are normalised log-probs
    m.backward()
    for got, want, what in zip(raw_grads(nvm, m), grads, ("encoder", "decoder", "lookup")):
        assert_close(got, want, 1e-4, f"{what} gradient")
    gd = m.get_grads(nvm.BLOCK_AE_DECODER)
    bd = gd[-(cfg.V + 1):]
    assert abs(bd.sum()) < 1e-4                                              # softmax-minus-onehot rows sum to zero
    m.adam_step()
    m.sync()
    m.close()


def test_literal_lookup_gradient_mode():
    """nvqa_set_lookup_grad_literal(1): the LookupTable only sees the weight-decay term (the literal reference, whose
    clones do not share gradWeight with the module parameters() exposes); encoder / decoder updates are unchanged."""
    nvm = nv()
    g = np.load(GOLD)
    ocfg = AE.AEConfig(**{k: int(g["cfg_" + k]) for k in ("V", "E", "H", "L", "T")})
    B = g["seq"].shape[0]
    e2, d2, l2 = g["enc"].copy(), g["dec"].copy(), g["lut"].copy()
    AE.train_step(ocfg, e2, d2, l2, [{}, {}, {}], g["seq"], lr=1e-3, seed=int(g["seed"]), literal_lookup_grad=True)
    m = make(nvm, nvm.AEConfig(V=ocfg.V, E=ocfg.E, H=ocfg.H, L=ocfg.L, T=ocfg.T, B=B), g["enc"], g["dec"], g["lut"], 3)
    m.set_lookup_grad_literal(True)
    m.set_batch_host(g["seq"], g["lengths"])
    m.forward(nvm.MODE_TRAIN, int(g["seed"]))
    m.backward()
    m.adam_step(lr=1e-3)
    for blk, key, want in ((nvm.BLOCK_AE_ENCODER, "enc", e2), (nvm.BLOCK_AE_DECODER, "dec", d2), (nvm.BLOCK_AE_LOOKUP, "lut", l2)):
        # the first Adam step is lr * g / (|g| + eps'): ill-conditioned element-wise where |g| is tiny, so the update is
        # checked in rel-L2 (1e-3) with a loose element-wise bound
        e2, em = rel_err(m.get_params(blk) - g[key], want - g[key])
        assert e2 <= 1e-3 and em <= 2e-2, f"{key} update (literal lookup gradient): rel-l2 {e2:.3e}, rel-max {em:.3e}"
    m.close()


@pytest.mark.parametrize("name,prec,tol", [("fp32_simt", 0, 1e-4), ("bf16x2", 3, 1e-4)])
@pytest.mark.parametrize("H", [128, 512])
def test_batch_size_and_tmax_changes_between_batches(name, prec, tol, H):
    """One handle, batches of different sizes and different longest sequences (encoder steps tmax, decoder tmax + 1;
    the decoder's slot 0 is the encoder's last slot): nothing of an earlier batch may leak into a later one.
    H = 512 runs the persistent kernels."""
    nvm = nv()
    cfg = nvm.AEConfig(V=203, E=H, H=H, L=1, T=8, B=96) if H == 512 else nvm.AEConfig(V=203, E=64, H=H, L=1, T=8, B=96)
    ocfg = AE.AEConfig(V=cfg.V, E=cfg.E, H=cfg.H, L=1, T=cfg.T)
    enc, dec, lut = nvm.synth_params_ae(cfg, seed=5)
    m = make(nvm, cfg, enc, dec, lut, prec)
    for i, (B, longest) in enumerate(((96, 8), (30, 3), (96, 8), (7, 8), (50, 5))):
        seq, lens = nvm.synth_batch_ae(cfg, B, seed=40 + i, min_len=1)
        seq[:, longest:] = 0
        lens = np.minimum(lens, longest).astype(np.int32)
        seq[0, :longest] = np.maximum(seq[0, :longest], 1)
        lens[0] = longest
        f, grads, ctx = AE.loss_and_grads(ocfg, enc, dec, lut, seq, seed=None, grad_clip=None, weight_decay=0)
        assert ctx["tmax"] == longest
        m.set_batch_host(seq, lens)
        m.forward(nvm.MODE_EVAL, 0)
        assert abs(m.loss() - f) <= tol * abs(f), f"batch {i} (B={B}, longest {longest}) loss {m.loss()} vs {f}"
        m.backward()
        for got, want, what in zip(raw_grads(nvm, m), grads, ("encoder", "decoder", "lookup")):
            assert_close(got, want, tol, f"batch {i} (B={B}, longest {longest}) {what} gradient")
    m.close()
