#!/bin/bash
# narrow last column tile (N = 208 for the two N = 200 GEMMs of layer 1): parity, then A/B against the round-1 shape rule
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_mirror_gpu.py tests/test_ae_gpu.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],4) for x in d['roofline']['classes']}
print('run $i', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'wgrad', c.get('lstm_wgrad_gemm'), 'dgrad', c.get('lstm_dgrad_gemm'))"
done
