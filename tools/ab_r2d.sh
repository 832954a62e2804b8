#!/bin/bash
out=gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2d_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
}
for c in 6 8 10 12 14 16 18 20; do run cap$c NVQA_AUX_CTAS=$c; done
run cap16b NVQA_AUX_CTAS=16
run cap12b NVQA_AUX_CTAS=12
