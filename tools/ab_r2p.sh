#!/bin/bash
# A/B: split-K reductions of the LSTM weight-gradient GEMMs on the side stream (default) vs inline
out=gpurun_out
for rep in 1 2; do
for d in 1 0; do
NVQA_DEFER_REDUCE=$d timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('defer_reduce=$d', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
