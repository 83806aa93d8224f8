#!/bin/bash
# fourth round-2 GPU pass: early prefetch (test + A/B), ncu of the fused forward kernel, reference timing on the free config,
# C++ host timing, 2k-iteration PSNR of both arms through the unmodified trainer
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02d}
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -k "prefetched or validate_mode or fused_forward" > $O/${TAG}_tests.log 2>&1
echo "--- prefetch / validate tests rc=$?"; tail -n 5 $O/${TAG}_tests.log | cut -c1-250
for EP in 1 0; do
  F2B_EARLY_PREFETCH=$EP timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $([ $EP = 0 ] && echo --no-ref-gpu) > $O/${TAG}_bench_ep$EP.json 2> $O/${TAG}_bench_ep$EP.err
  echo "--- bench early_prefetch=$EP rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_ep$EP.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:10]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "e2e_ms": d["e2e"]["ms_per_step"], "kernels": k, "forward_only": d.get("forward_only"),
                      "reference_gpu": (d.get("reference_gpu") or {}).get("ms_fwd_bwd_median"), "cpp_host": d.get("cpp_host")}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench_ep$EP.err").read()[-1500:])
PY
done
timeout 300 python scripts/timeline.py --steps 6 > $O/${TAG}_timeline.json 2> $O/${TAG}_timeline.err
python - <<PY
import json
d = json.load(open("$O/${TAG}_timeline.json"))
print("--- timeline", {k: v for k, v in d.items() if k != "streams"})
for s, v in d["streams"].items():
    print(" stream", s, round(v["busy_us"]), [(k["name"][:28], round(k["dur_us"]), round(k["overlapped_us"])) for k in v["by_kernel"][:6]])
PY
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:render_fwd_fused" -c 2 --launch-skip 2 -f -o $O/prof_${TAG}_fused \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_ncu_fused.log 2>&1
echo "--- ncu fused rc=$?"; ls -la $O/prof_${TAG}_fused.ncu-rep
timeout 600 python bench.py --config free --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_free.json 2> $O/${TAG}_bench_free.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench_free.json").read().strip().splitlines()[-1])
print("--- free", json.dumps({"ms_per_step": d["ms_per_step"], "reference_gpu": d.get("reference_gpu"), "cpp_host": d.get("cpp_host")})[:900])
PY
bash scripts/train_psnr.sh 2000 2>&1 | tail -n 3
cp gpurun_out/train_psnr.json $O/${TAG}_train_psnr_2k.json 2>/dev/null
tail -n 4 $O/train_ref.log | cut -c1-200; tail -n 4 $O/train_b200.log | cut -c1-200
du -sh $O
