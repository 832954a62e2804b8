#!/bin/bash
# autoencoder: one clearing launch per backward + deferred split-K reductions: parity, then config 5 (NVQA_AUX_STREAM / NVQA_PREZERO off = before)
timeout 600 python -m pytest tests/test_ae_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for cfg in "NVQA_X=1" "NVQA_PREZERO=0 NVQA_DEFER_REDUCE=0"; do
env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('extras') or {}
print('$cfg', round(d['value']), round(d['ms_per_step'],4), {k:(round(v['value']), v.get('ms_per_step')) for k,v in x.items() if isinstance(v,dict) and (k.startswith('ae') or k.startswith('arch2'))})"
done
