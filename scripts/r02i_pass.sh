#!/bin/bash
# ninth round-2 GPU pass: two-walker fused forward kernel (sanitizer, parity tests in both shapes, A/B), then the final build's bench
# lines for every config (reworked clock sampling / attempt selection)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02i}
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_render.py -q -x -p no:cacheprovider -k "fused_forward and free" > $O/${TAG}_san_ff.log 2>&1
echo "--- sanitizer fused forward (2 walkers) rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/${TAG}_san_ff.log | head -5
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_ref_parity.py tests/test_shim_dropin.py -m gpu -q -p no:cacheprovider > $O/${TAG}_tests_w2.log 2>&1
echo "--- render / ref / shim tests, 2 walkers rc=$?"; tail -n 3 $O/${TAG}_tests_w2.log | cut -c1-200
F2B_FUSED_WALKERS=1 timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -k "fused_forward or whole_image or validate" > $O/${TAG}_tests_w1.log 2>&1
echo "--- forward tests, 1 walker rc=$?"; tail -n 3 $O/${TAG}_tests_w1.log | cut -c1-200
for W in 2 1; do
  F2B_FUSED_WALKERS=$W timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_bench_w$W.json 2> $O/${TAG}_bench_w$W.err
  python -c "
import json; d=json.loads(open('$O/${TAG}_bench_w$W.json').read().strip().splitlines()[-1]); print('--- walkers $W', json.dumps({'ms_per_step': round(d['ms_per_step'],3), 'forward_only': d['forward_only']}))" || tail -n 5 $O/${TAG}_bench_w$W.err
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "--- bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:14]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e_ms": d["e2e"]["ms_per_step"], "attempts": [a["rejected"] for a in d["timing_attempts"]],
                      "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]], "roofline": d["roofline"], "kernels": k, "clocks": d["clocks"],
                      "forward_only": d.get("forward_only"), "reference_gpu": (d.get("reference_gpu") or {}).get("ms_fwd_bwd_median"), "cpp_host": (d.get("cpp_host") or {}).get("ms_fwd_bwd_median"),
                      "cpu_baseline": (d.get("cpu_baseline") or {}).get("value")}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench.err").read()[-1500:])
PY
for C in free nerf360 big20 big22; do
  timeout 600 python bench.py --config $C --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_$C.json 2> $O/${TAG}_bench_$C.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_$C.json").read().strip().splitlines()[-1])
    print("--- $C", json.dumps({"ms_per_step": round(d["ms_per_step"], 3), "value": round(d["value"]), "e2e": round(d["e2e"]["value"]), "attempts": [a["rejected"] for a in d["timing_attempts"]],
          "ref_ms": (d.get("reference_gpu") or {}).get("ms_fwd_bwd_median"), "cpp_ms": (d.get("cpp_host") or {}).get("ms_fwd_bwd_median"), "fwd_ms": (d.get("forward_only") or {}).get("ms_per_step")}))
except Exception as e:
    print("--- $C parse failed", e); print(open("$O/${TAG}_bench_$C.err").read()[-800:])
PY
done
du -sh $O
