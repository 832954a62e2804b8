#!/bin/bash
# arch 2 on the side stream (classifier weight gradient, deferred reductions, clearing launch): parity, then config 4
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_data_gpu.py -m gpu -x -q 2>&1 | tail -3
for aux in 1 0; do
NVQA_AUX_STREAM=$aux timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('extras') or {}
print('aux=$aux', round(d['value']), round(d['ms_per_step'],4), {k:(round(v['value']), v.get('ms_per_step')) for k,v in x.items() if isinstance(v,dict) and k.startswith('arch2')})"
done
