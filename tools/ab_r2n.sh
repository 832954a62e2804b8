#!/bin/bash
out=gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>$out/r2n_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches']/d['steps'])"
  tail -2 $out/r2n_$name.err
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run sideopt1 X=1
run sideopt0 NVQA_SIDE_OPT=0
