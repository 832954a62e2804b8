#!/bin/bash
out=gpurun_out
for ns in 64 0 16 32 128; do
NVQA_LSTM_POLL_NS=$ns timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('poll_ns=$ns', round(d['value']), round(d['ms_per_step'],4), 'fwd', c['lstm_recurrent_fwd'], 'bwd', c['lstm_recurrent_bwd'])"
done
