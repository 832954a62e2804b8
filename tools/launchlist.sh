#!/bin/bash
# ncu launch list of exactly the timed steps (per-launch rows kept, in launch order) -> gpurun_out/launches_$1.csv
tag=${1:-r2}
NVQA_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$tag.log 2>&1
tail -2 gpurun_out/ncu_launches_$tag.log; wc -l gpurun_out/launches_$tag.csv
