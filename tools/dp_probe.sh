#!/bin/bash
# N GPUs (first argument, default 2): the fused-exchange tests, then the exchange kernels alone for several chunk sizes / forms
N=${1:-2}
timeout 600 python -m pytest tests/test_dp_fused_gpu.py -m gpu -x -q 2>&1 | tail -4
for cfg in "NVQA_DP_CHUNK_KB=8" "NVQA_DP_CHUNK_KB=16" "NVQA_DP_CHUNK_KB=4" "NVQA_DP_CHUNK_KB=8 NVQA_DP_WHOLE=0" "NVQA_DP_CHUNK_KB=8 NVQA_DP_BULK_CTAS=74" "NVQA_DP_CHUNK_KB=16 NVQA_DP_BULK_CTAS=74"; do
  env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dp_probe.py 2>&1 | grep -E "dp_probe|Error|error" | head -5 | sed "s/^/$cfg /"
done
for cfg in "NVQA_X=1"; do
env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d.get('dp_check') or {}
print('$cfg bench n=$N', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'identical', c.get('replicas_identical'), 'vs_nccl', c.get('update_rel_l2_vs_nccl'))"
done
