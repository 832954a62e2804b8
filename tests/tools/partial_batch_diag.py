"""Diagnostic: per-row score error of eval batches smaller than cfg.B (small generic-path configuration)."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import novel_vqa_b200 as nvm
from novel_vqa_b200 import data
from oracle import arch1 as A

d = tempfile.mkdtemp()
js, qh5, ih5 = (os.path.join(d, n) for n in ("j.json", "q.h5", "i.h5"))
data.write_synthetic(js, qh5, ih5, n_train=96, n_val=40, n_test=50, n_img=12, T=12, V=60, O=10, I=40, seed=11)
ds = data.VqaDataset(js, qh5, ih5, splits=("test",), batch_size=32)
cfg = nvm.Arch1Config(V=60, E=24, H=64, L=2, I=40, C=48, O=10, T=12, B=32)
oc = A.Arch1Config(V=cfg.V, E=cfg.E, H=cfg.H, L=cfg.L, I=cfg.I, C=cfg.C, O=cfg.O, T=cfg.T, p=cfg.dropout)
enc, emb, mm = nvm.synth_params(cfg, seed=2)
enc, emb, mm = enc * 3, emb * 3, mm * 3
t = ds["test"]
for prec in (0, 3):
    m = nvm.Arch1Model(cfg, precision=prec)
    for blk, w in ((nvm.BLOCK_ENCODER, enc), (nvm.BLOCK_EMBEDDING, emb), (nvm.BLOCK_MULTIMODAL, mm)):
        m.set_params(blk, w)
    for order in ([(0, 32), (32, 50)], [(32, 50), (0, 32)], [(0, 18)], [(0, 32), (0, 18)]):
        for a, b in order:
            q, ln, fc7, _ = t.batch(np.arange(a + 1, b + 1))
            m.eval_step_host(q, ln, fc7)
            got = m.scores(b - a)
            _, _, want, ctx = A.jdj(oc, enc, emb, mm, q, ln, A.l2_normalize_rows(fc7), np.ones(b - a, np.int32), seed=None)
            err = np.abs(got - want).max(axis=1) / np.abs(want).max()
            st = m.state(b - a)
            serr = np.abs(st - ctx["tv_q"]).max(axis=1) / np.abs(ctx["tv_q"]).max()
            bad = np.nonzero(err > 1e-4)[0]
            print(f"prec {prec} rows [{a},{b}) after {order}: bad score rows {bad.tolist()} (max {err.max():.2e}); bad state rows "
                  f"{np.nonzero(serr > 1e-4)[0].tolist()} (max {serr.max():.2e}); lengths of bad rows {ln[bad].tolist()}")
    m.close()
