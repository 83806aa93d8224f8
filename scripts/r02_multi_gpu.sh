#!/bin/bash
# multi-GPU bench lines on ONE box (gpurun --gpus N): weak scaling (headline config, 4096 rays per GPU) and strong scaling
# (nerf360: one global 8192-ray batch split over the ranks), one process per GPU over NCCL, the driver's launch line.
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${TAG:-r02}
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
for C in wanjinyou nerf360; do
  for OV in 1 0; do
    [ $C = nerf360 ] && [ $OV = 0 ] && continue
    F2B_DP_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --config $C --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_n${N}_${C}_ov$OV.json 2> $O/${TAG}_n${N}_${C}_ov$OV.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_n${N}_${C}_ov$OV.json").read().strip().splitlines()[-1])
    print("--- N=$N $C overlap=$OV", json.dumps({"ms_per_step": round(d["ms_per_step"], 3), "value": round(d["value"]), "e2e": round(d["e2e"]["value"]), "scaling": d["scaling"],
          "rays_per_gpu": d["config"]["rays_per_gpu"], "n_gpus": d["n_gpus"]}))
except Exception as e:
    print("--- N=$N $C overlap=$OV failed", e); print(open("$O/${TAG}_n${N}_${C}_ov$OV.err").read()[-1500:])
PY
  done
done
if [ "${SINGLE:-1}" = "1" ]; then      # the N=1 line of the same box, for the ratio
  for C in wanjinyou nerf360; do
    timeout 600 python bench.py --config $C --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_n1_${C}.json 2> $O/${TAG}_n1_${C}.err
    python -c "
import json; d=json.loads(open('$O/${TAG}_n1_${C}.json').read().strip().splitlines()[-1]); print('--- N=1 $C', round(d['ms_per_step'],3), round(d['value']))"
  done
fi
