#!/bin/bash
# sanity of the final bench.py on 2 GPUs (collective branches of the clock-sample fallback / stall rejection) and on 1 GPU
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02m}
for C in wanjinyou nerf360; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --config $C --steps 20 --warmup 5 > $O/${TAG}_n2_$C.json 2> $O/${TAG}_n2_$C.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_n2_$C.json").read().strip().splitlines()[-1])
    print("--- N=2 $C", json.dumps({"ms_per_step": round(d["ms_per_step"], 3), "value": round(d["value"]), "e2e": round(d["e2e"]["value"]), "attempts": [a["rejected"] for a in d["timing_attempts"]],
          "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]], "clocks": d["clocks"]}))
except Exception as e:
    print("--- N=2 $C failed", e); print(open("$O/${TAG}_n2_$C.err").read()[-1500:])
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err
python -c "
import json; d=json.loads(open('$O/${TAG}_n1.json').read().strip().splitlines()[-1]); print('--- N=1', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks'], [a['rejected'] for a in d['timing_attempts']])"
