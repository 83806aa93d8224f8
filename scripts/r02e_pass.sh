#!/bin/bash
# fifth round-2 GPU pass: pre-sync host work + lean background march (tests, A/B bench lines, timeline), then the full 20k-iteration
# PSNR run of both arms through the unmodified trainer
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02e}
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_abi.py -m gpu -q -p no:cacheprovider > $O/${TAG}_tests.log 2>&1
echo "--- render tests rc=$?"; tail -n 5 $O/${TAG}_tests.log | cut -c1-250
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    print("--- bench $name", json.dumps({"ms_per_step": round(d["ms_per_step"], 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3)}))
except Exception as e:
    print("--- bench $name parse failed", e); print(open("$O/${TAG}_bench_$name.err").read()[-1200:])
PY
}
run default F2B_DUMMY=1
run no_presync F2B_PRESYNC_HOST=0
run no_bg F2B_MARCH_BG=0
run late_prefetch F2B_EARLY_PREFETCH=0
run default2 F2B_DUMMY=1
timeout 300 python scripts/timeline.py --steps 6 > $O/${TAG}_timeline.json 2> $O/${TAG}_timeline.err
python - <<PY
import json
d = json.load(open("$O/${TAG}_timeline.json"))
print("--- timeline", {k: v for k, v in d.items() if k != "streams"})
for s, v in d["streams"].items():
    print(" stream", s, round(v["busy_us"]), [(k["name"][:28], round(k["dur_us"]), round(k["overlapped_us"])) for k in v["by_kernel"][:7]])
PY
bash scripts/train_psnr.sh ${PSNR_ITERS:-20000} 2>&1 | tail -n 2
cp $O/train_psnr.json $O/${TAG}_train_psnr.json 2>/dev/null
grep -E "^[0-9]+: |Mean psnr" $O/train_ref.log | tr '\n' ' '; echo; grep -E "^[0-9]+: |Mean psnr" $O/train_b200.log | tr '\n' ' '; echo
du -sh $O
