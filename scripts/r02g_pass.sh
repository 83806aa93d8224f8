#!/bin/bash
# seventh round-2 GPU pass (N GPUs): data-parallel variants of the table-gradient exchange on the headline config (weak scaling),
# then the strong-scaling config with the default
cd "$(dirname "$0")/.."
N=${1:-4}; O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02g}
run() {  # name config env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_n${N}_$name.json 2> $O/${TAG}_n${N}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_n${N}_$name.json").read().strip().splitlines()[-1])
    print("--- N=$N $name", json.dumps({"ms_per_step": round(d["ms_per_step"], 3), "value": round(d["value"]), "e2e": round(d["e2e"]["value"]), "scaling": d["scaling"],
          "attempts": [a["rejected"] for a in d["timing_attempts"]], "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]]}))
except Exception as e:
    print("--- N=$N $name failed", e); print(open("$O/${TAG}_n${N}_$name.err").read()[-1500:])
PY
}
run weak_noverlap wanjinyou F2B_DP_OVERLAP=0
run weak_slabs4 wanjinyou F2B_DP_SLABS=4
run weak_slabs2 wanjinyou F2B_DP_SLABS=2
run strong_noverlap nerf360 F2B_DP_OVERLAP=0
run strong_slabs2 nerf360 F2B_DP_SLABS=2
