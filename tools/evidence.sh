#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, default bench (both arms), ncu launch list of the timed steps
# and one ncu --set full capture of the persistent LSTM kernels.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-final}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; tail -3 $out/pytest_$tag.log
timeout 300 python bench.py > $out/bench_${tag}_default.json 2> $out/bench_${tag}_default.err; tail -c 300 $out/bench_${tag}_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err
NVQA_LSTM_NOCOOP=1 NVQA_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_launches_$tag.log 2>&1
NVQA_LSTM_NOCOOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ --launch-skip 4 --launch-count 4 \
  -o $out/prof_lstm_$tag -f python tools/run_steps.py bf16x2 2 train > $out/ncu_full_$tag.log 2>&1
ls -la $out/prof_lstm_$tag.ncu-rep $out/launches_$tag.csv
