#!/bin/bash
# second round-2 GPU pass: new kernels (MLP backward recompute, fused forward) under the sanitizer + tests, the e2e stopwatch,
# A/B bench lines, then the whole suite with the reference binaries present.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02b}
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "recompute" > $O/${TAG}_san_rc.log 2>&1
echo "--- sanitizer recompute rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/${TAG}_san_rc.log | head -5
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_render.py -q -x -p no:cacheprovider -k "fused_forward and free" > $O/${TAG}_san_ff.log 2>&1
echo "--- sanitizer fused forward rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/${TAG}_san_ff.log | head -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -k "recompute or fused_forward or whole_image or train_step" > $O/${TAG}_new_tests.log 2>&1
echo "--- new tests rc=$?"; tail -n 12 $O/${TAG}_new_tests.log | cut -c1-250
for V in "1 1" "0 1" "1 0"; do set -- $V
  F2B_FUSED_LAUNCH=$2 timeout 300 python scripts/e2e_probe.py $([ $1 = 0 ] && echo --no-pipeline-march) > $O/${TAG}_e2e_p$1_f$2.json 2> $O/${TAG}_e2e_p$1_f$2.err
  echo "--- e2e probe pipeline=$1 fused_launch=$2"; python -c "import json;d=json.load(open('$O/${TAG}_e2e_p$1_f$2.json'));print(d['median_ms'])" || tail -n 5 $O/${TAG}_e2e_p$1_f$2.err
done
timeout 300 python scripts/mlp_probe.py > $O/${TAG}_mlp_probe.json 2> $O/${TAG}_mlp_probe.err; echo "--- mlp probe"; cat $O/${TAG}_mlp_probe.json
for RC in 1 0; do
  F2B_MLP_RECOMPUTE=$RC timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $([ $RC = 0 ] && echo --no-ref-gpu) > $O/${TAG}_bench_rc$RC.json 2> $O/${TAG}_bench_rc$RC.err
  echo "--- bench recompute=$RC rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_rc$RC.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:12]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "e2e_ms": d["e2e"]["ms_per_step"], "kernels": k, "forward_only": d.get("forward_only"),
                      "reference_gpu": d.get("reference_gpu")}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench_rc$RC.err").read()[-1500:])
PY
done
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $O/${TAG}_pytest_gpu.log 2>&1
echo "--- pytest -m gpu (all): rc=$?"; tail -n 15 $O/${TAG}_pytest_gpu.log | cut -c1-250
du -sh $O
