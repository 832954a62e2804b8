"""A few full-size text-autoencoder training steps (BASELINE config 5): the ncu target for arch 3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import novel_vqa_b200 as nv  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = nv.AEConfig()
m = nv.AEModel(cfg, precision=nv.PREC_BF16X2)
for blk, w in zip((0, 1, 2), nv.synth_params_ae(cfg, seed=123)):
    m.set_params(blk, w)
seq, lens = nv.synth_batch_ae(cfg, cfg.B, seed=123)
m.set_batch_host(seq, lens)
for i in range(steps):
    m.forward(nv.MODE_TRAIN, i)
    m.backward()
    m.adam_step()
m.sync()
print("loss", m.loss(), "launches", nv.launch_count())
