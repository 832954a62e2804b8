#!/bin/bash
out=gpurun_out
for mode in nccl fused; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --dp $mode > $out/bench_n2_$mode.json 2> $out/bench_n2_$mode.err
tail -2 $out/bench_n2_$mode.err | cut -c1-200
python - <<PY
import json
d=json.loads(open("$out/bench_n2_$mode.json").read().strip().splitlines()[-1])
c={x['kernel']:round(x['ms_per_step'],3) for x in d['roofline']['classes']}
print("$mode", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "opt class", c.get("clamp_rmsprop"))
PY
done
