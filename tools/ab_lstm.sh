#!/bin/bash
# A/B of the persistent LSTM kernel variants: NVQA_LSTM_FWD_SPLIT x NVQA_LSTM_STACK (bench classes, ms per step)
for cfg in ${CFGS:-"1 0" "0 0" "1 1" "0 1"}; do
  set -- $cfg
  export NVQA_LSTM_FWD_SPLIT=$1 NVQA_LSTM_STACK=$2
  if [ "$2" = "1" ]; then timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_size_config1 or ragged or golden" 2>&1 | tail -1; fi
  timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_$1$2.json 2> gpurun_out/ab_$1$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$1$2.json").read().strip().splitlines()[-1])
c={x["kernel"]:round(x["ms_per_step"],3) for x in d["roofline"]["classes"]}
print("split=$1 stack=$2  value %.0f  ms/step %.3f  fwd %.3f  bwd %.3f" % (d["value"], d["ms_per_step"], c["lstm_recurrent_fwd"], c["lstm_recurrent_bwd"]))
PY
done
