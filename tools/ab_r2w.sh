#!/bin/bash
# loss reduction off the critical path + reduction-aware split-K rule; then the GPU suite
run() {
env $1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('extras') or {}
def g(k):
    v=x.get(k,{})
    return round(v.get('value',0)) if isinstance(v,dict) else v
c={x['kernel']:round(x['ms_per_step'],4) for x in d['roofline']['classes']}
print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'head', c.get('head_fwd_gemm'), c.get('head_bwd_gemm'), {k:g(k) for k in x})"
}
run "NVQA_X=0" ""
run "NVQA_X=0" "--no-extras"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
