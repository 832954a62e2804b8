#!/bin/bash
# 8 GPUs: the fused-exchange test (small), then the weak-scaling bench at 8 and 4 ranks
timeout 300 python -m pytest tests/test_dp_fused_gpu.py -m gpu -x -q -k small 2>&1 | tail -3
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/dp8b_n$N.err | tee gpurun_out/bench_r2g_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d.get('dp_check') or {}
print('bench n=$N', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'identical', c.get('replicas_identical'), 'vs_nccl', c.get('update_rel_l2_vs_nccl'))"
done
