#!/bin/bash
# A/B of the persistent (double-buffered TMEM) GEMM engine against one tile per CTA
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_ae_gpu.py -m gpu -x -q 2>&1 | tail -2
for pz in 1 0 1 0; do
  NVQA_GEMM_PERSIST=$pz timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('persist=$pz', round(d['value']), round(d['ms_per_step'],3), c, {k:(round(v['ms_per_step'],3) if v.get('ms_per_step') else round(v['value'])) for k,v in d['extras'].items()})"
done
