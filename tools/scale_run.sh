#!/bin/bash
# usage: tools/scale_run.sh N [fused|nccl]  -> gpurun_out/bench_n${N}_${dp}.json
N=$1; DP=${2:-fused}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 5 --dp $DP --no-extras > gpurun_out/bench_n${N}_${DP}.json 2> gpurun_out/bench_n${N}_${DP}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n${N}_${DP}.json").read().strip().splitlines()[-1])
print("N=${N} ${DP}: value %.0f  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("   ", {c["kernel"]: round(c["ms_per_step"], 3) for c in d["roofline"]["classes"]})
PY
