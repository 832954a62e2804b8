#!/bin/bash
# where does the pair kernel's epilogue time go?  (timing experiment: results invalid under dbg != 0)
# NVQA_GEMM_EPI_DBG existed only in the working tree of this experiment (bit 1: no global stores, 2: no shared-memory
# staging either, 4: first 32-column chunk only); the numbers are in DESIGN 5.1, the rewrite that followed is ab_r2t.sh
run() {
NVQA_GEMM_EPI_DBG=$1 timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],4) for x in d['roofline']['classes']}
print('epi_dbg=$1', round(d['ms_per_step'],4), 'wgrad', c.get('lstm_wgrad_gemm'), 'dgrad', c.get('lstm_dgrad_gemm'), 'inproj', c.get('lstm_inproj_gemm'))"
}
run 0
run 1
run 2
run 4
run 0
