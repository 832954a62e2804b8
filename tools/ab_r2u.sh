#!/bin/bash
# programmatic dependent launch (NVQA_PDL): parity first, then A/B
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() {
NVQA_PDL=$1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('extras') or {}
def g(k):
    v=x.get(k,{})
    return round(v.get('value',0)) if isinstance(v,dict) else v
print('pdl=$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), {k:g(k) for k in x})"
}
run 1
run 0
run 1
run 0
