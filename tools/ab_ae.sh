#!/bin/bash
timeout 900 python -m pytest tests/test_ae_gpu.py -m gpu -x -q 2>&1 | tail -3
for f in 1 0; do
NVQA_AE_FUSED=$f timeout 200 python tools/profile_ae.py 2>&1 | tail -14
done
