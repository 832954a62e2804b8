#!/bin/bash
# Timing experiments on the persistent LSTM forward kernel (NVQA_LSTM_DEBUG bits, see lstm_persistent_v2.cu).
# Prints the last timeline (layer 2) for every variant.  Results of variants with bits 2..32 / 128 are numerically wrong on purpose.
for split in ${SPLITS:-0 1}; do
  for dbg in ${DBGS:-7 23 39}; do
    echo "== split=$split dbg=$dbg"
    NVQA_LSTM_FWD_SPLIT=$split NVQA_LSTM_DEBUG=$dbg timeout 60 python tools/run_steps.py bf16x2 1 fwd 2>&1 | grep -A7 "lstm_fwd_v2 timeline" | tail -7
  done
done
