#!/bin/bash
# tenth round-2 GPU pass (8 GPUs): the weak- and strong-scaling lines at N = 8 with the round's defaults
cd "$(dirname "$0")/.."
N=${1:-8}; O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02j}
nvidia-smi -L | wc -l
for C in wanjinyou nerf360; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --config $C --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_n${N}_$C.json 2> $O/${TAG}_n${N}_$C.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_n${N}_$C.json").read().strip().splitlines()[-1])
    print("--- N=$N $C", json.dumps({"ms_per_step": round(d["ms_per_step"], 3), "value": round(d["value"]), "e2e": round(d["e2e"]["value"]), "scaling": d["scaling"],
          "rays_per_gpu": d["config"]["rays_per_gpu"], "attempts": [a["rejected"] for a in d["timing_attempts"]], "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]],
          "clocks": d["clocks"]}))
except Exception as e:
    print("--- N=$N $C failed", e); print(open("$O/${TAG}_n${N}_$C.err").read()[-1500:])
PY
done
