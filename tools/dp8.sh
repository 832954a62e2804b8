#!/bin/bash
# 8-GPU box: fused data-parallel correctness at world 8 (small config) + bench at N = 8 (overlap on / off), 4, 2
out=gpurun_out
timeout 300 python -m pytest tests/test_dp_fused_gpu.py -m gpu -x -q -k small 2>&1 | tail -3
runn() {
  n=$1; name=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_$name.json 2> $out/bench_$name.err
  python - <<PY
import json
d=json.loads(open("$out/bench_$name.json").read().strip().splitlines()[-1])
c=d.get("dp_check") or {}
print("$name", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "identical", c.get("replicas_identical"), "vs_nccl", c.get("update_rel_l2_vs_nccl"))
PY
}
runn 8 n8 X=1
runn 8 n8_noov NVQA_DP_OVERLAP=0
runn 4 n4 X=1
runn 2 n2 X=1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', round(d['value']), d['ms_per_step'])"
