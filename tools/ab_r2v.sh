#!/bin/bash
# one clearing launch per backward (NVQA_PREZERO) and programmatic dependent launch of the recurrent kernels (NVQA_LSTM_PDL)
run() {
env $1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('extras') or {}
def g(k):
    v=x.get(k,{})
    return round(v.get('value',0)) if isinstance(v,dict) else v
c={x['kernel']:round(x['ms_per_step'],4) for x in d['roofline']['classes']}
print('$1', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'fwd', c.get('lstm_recurrent_fwd'), 'bwd', c.get('lstm_recurrent_bwd'), {k:g(k) for k in x})"
}
run "NVQA_X=0" ""
run "NVQA_LSTM_PDL=0" "--no-extras"
run "NVQA_LSTM_PDL=0 NVQA_PREZERO=0" "--no-extras"
run "NVQA_X=0" "--no-extras"
