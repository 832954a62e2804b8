"""Generates tests/golden/ae_small.npz from the text-autoencoder oracle (oracle/ae.py).  The reference (Torch7 Lua)
cannot run in this image, so these vectors pin the oracle + CUDA path against regressions, not against Torch7
("parity unpinned").   python tests/golden/make_golden_ae.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import ae as AE  # noqa: E402

CFG = dict(V=150, E=32, H=64, L=1, T=7)
B, SEED = 24, 42


def main():
    cfg = AE.AEConfig(**CFG)
    r = np.random.default_rng(2024)
    enc = r.uniform(-0.08, 0.08, cfg.n_enc).astype(np.float32)
    dec = r.uniform(-0.08, 0.08, cfg.n_dec).astype(np.float32)
    lut = r.uniform(-0.08, 0.08, cfg.n_lut).astype(np.float32)
    lens = r.integers(1, cfg.T + 1, B)
    lens[3] = cfg.T
    seq = np.zeros((B, cfg.T), dtype=np.int32)
    for b in range(B):
        seq[b, :lens[b]] = r.integers(1, cfg.V + 1, lens[b])
    f, grads, ctx = AE.loss_and_grads(cfg, enc, dec, lut, seq, seed=SEED, keep_logprobs=True)
    # three optimizer steps of the training loop
    e2, d2, l2 = enc.copy(), dec.copy(), lut.copy()
    states = [{}, {}, {}]
    losses = [AE.train_step(cfg, e2, d2, l2, states, seq, lr=1e-3, seed=SEED + i)[0] for i in range(3)]
    out = dict(enc=enc, dec=dec, lut=lut, seq=seq, lengths=lens.astype(np.int32), seed=SEED, loss=np.float64(f),
               g_enc=grads[0], g_dec=grads[1], g_lut=grads[2], targets=ctx["targets"], n=ctx["n"],
               logprobs1=ctx["logprobs"][1], enc_final=ctx["enc_final"], traj_losses=np.array(losses),
               traj_enc=e2, traj_dec=d2, traj_lut=l2)
    out.update({"cfg_" + k: v for k, v in CFG.items()})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ae_small.npz"), **out)
    print("loss", f, "n", ctx["n"], "tmax", ctx["tmax"], "traj", losses)


if __name__ == "__main__":
    main()
