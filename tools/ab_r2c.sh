#!/bin/bash
# round 2: side-stream image branch / AxB weight gradients beside the recurrent kernels; with and without the cooperative attribute
out=gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2c_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']/d['steps'])"
  tail -2 $out/r2c_$name.err
}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_mirror_gpu.py tests/test_data_gpu.py -m gpu -x -q 2>&1 | tail -3
run aux1 X=1
run aux0 NVQA_AUX_STREAM=0
run aux1_nocoop NVQA_LSTM_NOCOOP=1
run aux0_nocoop NVQA_AUX_STREAM=0 NVQA_LSTM_NOCOOP=1
run aux1_cap16 NVQA_AUX_CTAS=16
