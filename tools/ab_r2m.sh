#!/bin/bash
out=gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2m_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('$name', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'fwd', c['lstm_recurrent_fwd'], 'bwd', c['lstm_recurrent_bwd'])"
  tail -2 $out/r2m_$name.err
}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_ae_gpu.py -m gpu -x -q 2>&1 | tail -3
run w6 X=1
run w0 NVQA_LSTM_W1TMEM=0
run w4 NVQA_LSTM_W1TMEM=4
