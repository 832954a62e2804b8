"""Worker of tests/test_variants_gpu.py: two full-size arch1 training steps (BASELINE config 1, bf16x2) under whatever
NVQA_* switches the environment carries; prints the loss trajectory, the gradients of the second step (as float32 bytes in a
temporary .npz) so that the parent can compare kernel variants against the default build."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import novel_vqa_b200 as nv  # noqa: E402

out = sys.argv[1]
cfg = nv.Arch1Config()
m = nv.Arch1Model(cfg, precision=nv.PREC_BF16X2)
enc, emb, mm = nv.synth_params(cfg, seed=123)
for blk, w in ((nv.BLOCK_ENCODER, enc), (nv.BLOCK_EMBEDDING, emb), (nv.BLOCK_MULTIMODAL, mm)):
    m.set_params(blk, w)
q, ln, fc7, lab = nv.synth_batch(cfg, 500, seed=321, min_len=3)
losses = []
for i in range(2):
    losses.append(m.train_step_host(q, ln, fc7, lab, 3e-4, 77 + i))
m.set_batch_host(q, ln, fc7, lab)
m.forward(nv.MODE_TRAIN, 99)
scores = m.scores(500)
m.backward()
np.savez(out, losses=np.array(losses), scores=scores, genc=m.get_grads(nv.BLOCK_ENCODER), gemb=m.get_grads(nv.BLOCK_EMBEDDING),
         gmm=m.get_grads(nv.BLOCK_MULTIMODAL), penc=m.get_params(nv.BLOCK_ENCODER), pmm=m.get_params(nv.BLOCK_MULTIMODAL))
m.close()
print("VARIANT_OK")
