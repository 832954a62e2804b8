"""Runs a few full-size (BASELINE config 1) training steps; used as the ncu target.
  python tools/run_steps.py [precision] [steps] [fwd|train]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import novel_vqa_b200 as nv  # noqa: E402

prec = {"fp32_simt": 0, "bf16x3": 1, "bf16": 2, "bf16x2": 3}[sys.argv[1] if len(sys.argv) > 1 else "bf16x2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
what = sys.argv[3] if len(sys.argv) > 3 else "train"
cfg = nv.Arch1Config()
m = nv.Arch1Model(cfg, precision=prec)
enc, emb, mm = nv.synth_params(cfg)
for blk, w in ((nv.BLOCK_ENCODER, enc), (nv.BLOCK_EMBEDDING, emb), (nv.BLOCK_MULTIMODAL, mm)):
    m.set_params(blk, w)
q, ln, fc7, lab = nv.synth_batch(cfg, 500)
m.set_batch_host(q, ln, fc7, lab)
for i in range(steps):
    m.forward(nv.MODE_TRAIN, i)
    if what == "train":
        m.backward()
        m.rmsprop_step(3e-4)
m.sync()
print("loss", m.loss(), "launches", nv.launch_count())
