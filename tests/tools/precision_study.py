"""Prints the end-to-end error of every precision mode of the CUDA path against the fp32/fp64 CPU oracle on
BASELINE config 1 (B=500).  Run on the GPU box:  python tests/tools/precision_study.py [--f64]
(lives under tests/: it calls the oracle, which only test code may do)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import novel_vqa_b200 as nv  # noqa: E402
from oracle import arch1 as A  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / np.linalg.norm(b), np.abs(a - b).max() / np.abs(b).max()


def main():
    cfg = nv.Arch1Config()
    oc = A.Arch1Config()
    enc, emb, mm = nv.synth_params(cfg, seed=123)
    q, ln, fc7, lab = nv.synth_batch(cfg, 500, seed=123)
    dt = np.float64 if "--f64" in sys.argv else np.float32
    t0 = time.time()
    f, grads, scores, ctx = A.jdj(oc, enc, emb, mm, q, ln, A.l2_normalize_rows(fc7.astype(dt)), lab, seed=7, dtype=dt)
    print(f"oracle {dt.__name__}: loss {f:.6f} ({time.time() - t0:.1f}s)")
    for name, prec in (("fp32_simt", 0), ("bf16x3", 1), ("bf16x2", 3), ("bf16", 2)):
        m = nv.Arch1Model(cfg, precision=prec)
        for blk, w in ((nv.BLOCK_ENCODER, enc), (nv.BLOCK_EMBEDDING, emb), (nv.BLOCK_MULTIMODAL, mm)):
            m.set_params(blk, w)
        m.set_batch_host(q, ln, fc7, lab)
        m.forward(nv.MODE_TRAIN, 7)
        m.backward()
        out = [f"{name:10s} loss {abs(m.loss() - f) / abs(f):.2e}", "scores %.2e/%.2e" % rel(m.scores(500), scores),
               "state %.2e/%.2e" % rel(m.state(500), ctx["tv_q"])]
        for blk, gw, nm in zip((0, 1, 2), grads, ("genc", "gemb", "gmm")):
            out.append(nm + " %.2e/%.2e" % rel(np.clip(m.get_grads(blk), -10, 10), gw))
        print("  ".join(out), flush=True)
        m.close()


if __name__ == "__main__":
    main()
