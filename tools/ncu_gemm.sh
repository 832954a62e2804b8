#!/bin/bash
# ncu --set full of the batched tcgen05 GEMMs of one training step (second step: launches 0..N of step 2)
tag=${1:-r2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm --launch-skip 16 --launch-count 16 \
  -o gpurun_out/prof_gemm_$tag -f python tools/run_steps.py bf16x2 2 train > gpurun_out/ncu_gemm_$tag.log 2>&1
tail -3 gpurun_out/ncu_gemm_$tag.log; ls -la gpurun_out/prof_gemm_$tag.ncu-rep
