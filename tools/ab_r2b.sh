#!/bin/bash
# round 2: 4-D TMA boxes in the backward kernel (default on) vs off; parity subset first
out=gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2b_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('$name', round(d['value']), round(d['ms_per_step'],3), c)"
  tail -2 $out/r2b_$name.err
}
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_ae_gpu.py -m gpu -x -q 2>&1 | tail -3
run box4 X=1
run nobox4 NVQA_LSTM_BOX4D=0
