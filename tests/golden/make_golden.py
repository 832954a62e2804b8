"""Generates tests/golden/arch1_small.npz from the CPU oracle (oracle/arch1.py).

The reference ships no golden vectors and Torch7 cannot run here, so the fixtures
pin the ORACLE's results (fp64 truth + fp32) for a small arch1 configuration with variable
question lengths; tests/test_oracle.py re-derives them, tests/test_parity_gpu.py checks the CUDA
path against them.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import arch1 as A  # noqa: E402

CFG = dict(V=50, E=8, H=12, L=2, I=16, C=20, O=11, T=7)
B = 6
SEED_DATA, SEED_DROP = 2024, 77


def make_inputs():
    cfg = A.Arch1Config(**CFG)
    r = np.random.default_rng(SEED_DATA)
    lengths = np.array([7, 3, 7, 1, 5, 3], dtype=np.int32)
    q = np.zeros((B, cfg.T), dtype=np.int32)
    for b in range(B):
        q[b, :lengths[b]] = r.integers(1, cfg.V + 1, lengths[b])
    enc = r.uniform(-0.3, 0.3, cfg.n_enc).astype(np.float32)
    emb = r.uniform(-0.3, 0.3, cfg.n_emb).astype(np.float32)
    mm = r.uniform(-0.3, 0.3, cfg.n_mm).astype(np.float32)
    fc7 = np.maximum(0, r.standard_normal((B, cfg.I))).astype(np.float32)
    labels = r.integers(1, cfg.O + 1, B).astype(np.int32)
    return cfg, q, lengths, enc, emb, mm, fc7, labels


def compute():
    cfg, q, lengths, enc, emb, mm, fc7, labels = make_inputs()
    q_ra = A.right_align(q, lengths)
    words, sizes, sidx, inv = A.sort_encoding_right_align(q_ra, lengths)
    out = dict(q=q, lengths=lengths, q_ra=q_ra, words=words, batch_sizes=sizes, sort_index=sidx + 1,
               sort_index_inverse=inv + 1, enc=enc, emb=emb, mm=mm, fc7=fc7, labels=labels)
    for tag, seed in (("eval", None), ("train", SEED_DROP)):
        for dt, dn in ((np.float64, "f64"), (np.float32, "f32")):
            fv = A.l2_normalize_rows(fc7.astype(dt))
            f, g, scores, ctx = A.jdj(cfg, enc, emb, mm, q_ra, lengths, fv, labels, seed=seed, dtype=dt)
            out[f"{tag}_{dn}_loss"] = np.array(f)
            out[f"{tag}_{dn}_scores"] = scores
            out[f"{tag}_{dn}_state"] = ctx["tv_q"]
            out[f"{tag}_{dn}_genc"], out[f"{tag}_{dn}_gemb"], out[f"{tag}_{dn}_gmm"] = g
            if tag == "eval":
                out[f"eval_{dn}_argmax"] = A.argmax_first(scores)
    # three RMSprop iterations in training mode (fp32), seeds SEED_DROP + it
    w = [enc.copy(), emb.copy(), mm.copy()]
    ms = [np.zeros_like(x) for x in w]
    lr, losses = 3e-4, []
    fv = A.l2_normalize_rows(fc7)
    for it in range(3):
        f, lr = A.train_step(cfg, w[0], w[1], w[2], ms, (q_ra, lengths, fv, labels), lr, seed=SEED_DROP + it)
        losses.append(f)
    out["traj_losses"] = np.array(losses, dtype=np.float32)
    out["traj_enc"], out["traj_emb"], out["traj_mm"] = w
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arch1_small.npz")
    np.savez_compressed(path, **compute())
    print("wrote", path, os.path.getsize(path), "bytes")
