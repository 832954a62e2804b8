#!/bin/bash
# TMA-store epilogue of the pair GEMM kernel: parity first, then A/B (NVQA_GEMM_TMA_STORE)
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_ae_gpu.py tests/test_mirror_gpu.py -m gpu -x -q 2>&1 | tail -5
run() {
NVQA_GEMM_TMA_STORE=$1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],4) for x in d['roofline']['classes']}
x=d.get('extras') or {}
def g(k):
    v=x.get(k,{})
    return round(v.get('value',0)) if isinstance(v,dict) else v
print('tma_store=$1', round(d['value']), round(d['ms_per_step'],4), 'wgrad', c.get('lstm_wgrad_gemm'), 'dgrad', c.get('lstm_dgrad_gemm'), 'inproj', c.get('lstm_inproj_gemm'), {k:g(k) for k in x})"
}
run 1
run 0
run 1
run 0
