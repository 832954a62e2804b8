#!/bin/bash
# autoencoder kernels under programmatic dependent launch: parity, then the extras lines
timeout 600 python -m pytest tests/test_ae_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for pdl in 1 0; do
NVQA_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('extras') or {}
print('pdl=$pdl', round(d['value']), round(d['ms_per_step'],4), {k:(round(v['value']), v.get('ms_per_step')) for k,v in x.items() if isinstance(v,dict)})"
done
