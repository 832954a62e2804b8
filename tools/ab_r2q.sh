#!/bin/bash
# A/B: tile shape / split-K count of the LSTM dgrad and wgrad GEMMs (NVQA_GEMM_OVERRIDE)
run() {
NVQA_GEMM_OVERRIDE="$1" timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],4) for x in d['roofline']['classes']}
print('override=[$1]', round(d['value']), round(d['ms_per_step'],4), 'wgrad', c.get('lstm_wgrad_gemm'), 'dgrad', c.get('lstm_dgrad_gemm'), 'inproj', c.get('lstm_inproj_gemm'))"
}
run ""
run "13000x512x2048:128:1"
run "2048x512x13000:256:9"
run "2048x512x13000:256:9,2048x200x13000:256:9"
run "13000x512x2048:128:1,2048x512x13000:256:9,2048x200x13000:256:9"
run "2048x512x13000:128:2"
run "13000x2048x512:128:1"
run ""
