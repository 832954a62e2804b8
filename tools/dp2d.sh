#!/bin/bash
# 2 GPUs, final build: the fused-exchange tests and the weak-scaling line
timeout 600 python -m pytest tests/test_dp_fused_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tee gpurun_out/bench_r2i_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d.get('dp_check') or {}
print('bench n=2', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'identical', c.get('replicas_identical'), 'vs_nccl', c.get('update_rel_l2_vs_nccl'))"
