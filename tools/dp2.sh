#!/bin/bash
# 2-GPU correctness + bench of the fused data-parallel step
out=gpurun_out
timeout 600 python -m pytest tests/test_dp_fused_gpu.py -m gpu -x -q 2>&1 | tail -5
for ov in 1 0; do
NVQA_DP_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_ov$ov.json 2> $out/bench_n2_ov$ov.err
tail -3 $out/bench_n2_ov$ov.err
python - <<PY
import json
d=json.loads(open("$out/bench_n2_ov$ov.json").read().strip().splitlines()[-1])
print("overlap=$ov", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("dp_check"))
PY
done
