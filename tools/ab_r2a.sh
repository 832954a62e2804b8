#!/bin/bash
# round 2, first GPU call: baseline state + the two variants written at the end of round 1 that never ran (BOX4D / POLL1)
out=gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2a_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('$name', round(d['value']), round(d['ms_per_step'],3), c)"
  tail -2 $out/r2a_$name.err
}
run base X=1
for v in "NVQA_LSTM_BOX4D=1" "NVQA_LSTM_POLL1=1" "NVQA_LSTM_BOX4D=1 NVQA_LSTM_POLL1=1"; do
  tag=$(echo $v | tr ' =' '__')
  echo "== $v"
  env $v timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_size_config1 or batch_size_change or golden or edge" 2>&1 | tail -3
  run $tag $v
done
