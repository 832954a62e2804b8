#!/bin/bash
# cta_group::2 pair variant of the forward recurrent kernel: parity subset + bench A/B
export NVQA_LSTM_PAIR=1
timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_size_config1 or batch_size_change or golden" 2>&1 | tail -6
for pz in 1 0; do
  NVQA_LSTM_PAIR=$pz timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/ab_pair$pz.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print('pair=$pz', round(d['value']), round(d['ms_per_step'],3), c)"
  tail -3 gpurun_out/ab_pair$pz.err
done
