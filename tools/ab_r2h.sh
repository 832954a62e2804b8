#!/bin/bash
out=gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2h_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('$name', round(d['value']), round(d['ms_per_step'],4), 'bwd', c['lstm_recurrent_bwd'], 'pw_bwd', c['pointwise_bwd'])"
  tail -2 $out/r2h_$name.err
}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_size_config1 or batch_size_change or golden or edge" 2>&1 | tail -3
run bpair1 X=1
run bpair0 NVQA_LSTM_BWD_PAIR=0
run bpair1_noaux NVQA_AUX_STREAM=0
run bpair0_noaux NVQA_LSTM_BWD_PAIR=0 NVQA_AUX_STREAM=0
