"""Per-kernel-class timing of the text-autoencoder step (BASELINE config 5) through nvqa_profile."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import novel_vqa_b200 as nv  # noqa: E402

cfg = nv.AEConfig()
m = nv.AEModel(cfg, precision=nv.PREC_BF16X2)
for blk, w in zip((0, 1, 2), nv.synth_params_ae(cfg, seed=123)):
    m.set_params(blk, w)
seq, lens = nv.synth_batch_ae(cfg, cfg.B, seed=123)
m.set_batch_host(seq, lens)
for i in range(3):
    m.forward(nv.MODE_TRAIN, i); m.backward(); m.adam_step()
nv._lib.check(m.lib.nvqa_profile(m.handle, 1))
n = 5
for i in range(n):
    m.forward(nv.MODE_TRAIN, 10 + i); m.backward(); m.adam_step()
buf = ctypes.create_string_buffer(8192)
nv._lib.check(m.lib.nvqa_profile_report(m.handle, buf, 8192))
tot = 0
for c in json.loads(buf.value.decode()):
    if c["launches"]:
        print(f'{c["kernel"]:22s} {c["ms"] / n:8.3f} ms/step  {c["launches"] / n:5.1f} launches')
        tot += c["ms"] / n
print("sum", round(tot, 3))
