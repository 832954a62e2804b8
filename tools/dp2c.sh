#!/bin/bash
out=gpurun_out
for ctas in 1184 592 296 148; do
NVQA_DP_OVERLAP=0 NVQA_DP_MAIN_CTAS=$ctas timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > $out/bench_n2_c$ctas.json 2> $out/bench_n2_c$ctas.err
python - <<PY
import json
d=json.loads(open("$out/bench_n2_c$ctas.json").read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print("ctas $ctas", round(d["value"]), round(d["ms_per_step"],4), "opt class", c.get("clamp_rmsprop"))
PY
done
