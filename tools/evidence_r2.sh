#!/bin/bash
# Round-2 evidence on one B200 (outputs under gpurun_out/, tag = $1): GPU suite, both bench arms, ncu launch list of the
# timed steps, ncu --set full of the recurrent kernels and of the batched GEMMs, smoke() under ncu (what the driver does).
tag=${1:-r2}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > $out/pytest_$tag.log 2>&1; tail -4 $out/pytest_$tag.log
timeout 600 python bench.py > $out/bench_${tag}_default.json 2> $out/bench_${tag}_default.err; tail -c 200 $out/bench_${tag}_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err
bash tools/launchlist.sh $tag
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ --launch-skip 4 --launch-count 4 \
  -o $out/prof_lstm_$tag -f python tools/run_steps.py bf16x2 2 train > $out/ncu_full_$tag.log 2>&1
bash tools/ncu_gemm.sh $tag
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/smoke_launches_$tag.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_ncu_$tag.log 2>&1; tail -2 $out/smoke_ncu_$tag.log
grep -c "lstm_fwd_v2_kernel\|lstm_bwd_v3_kernel" $out/smoke_launches_$tag.csv
ls -la $out/prof_lstm_$tag.ncu-rep $out/prof_gemm_$tag.ncu-rep $out/launches_$tag.csv
