#!/bin/bash
# final round-2 GPU pass on the final commit: whole suite, smoke, bench line, ncu of the fused forward kernel
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02k}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $O/${TAG}_pytest_gpu.log 2>&1
echo "--- pytest -m gpu (all): rc=$?"; tail -n 4 $O/${TAG}_pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "--- smoke rc=$?"; tail -n 1 $O/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "--- bench (no flags) rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:12]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e": d["e2e"]["value"], "e2e_ms": d["e2e"]["ms_per_step"], "attempts": [a["rejected"] for a in d["timing_attempts"]],
                      "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]], "roofline_frac": d["roofline"]["frac"], "governing": d["roofline"]["governing"], "kernels": k,
                      "clocks": d["clocks"], "forward_only": d.get("forward_only"), "reference_gpu": (d.get("reference_gpu") or {}).get("ms_fwd_bwd_median"),
                      "ref_validate": (d.get("reference_gpu") or {}).get("ms_validate_median"), "cpp_host": (d.get("cpp_host") or {}).get("ms_fwd_bwd_median"),
                      "cpu_baseline": (d.get("cpu_baseline") or {}).get("value"), "launches": d["gpu_launches"]}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench.err").read()[-1500:])
PY
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:render_fwd_fused" -c 2 --launch-skip 2 -f -o $O/prof_${TAG}_fused \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_ncu_fused.log 2>&1
echo "--- ncu fused rc=$?"; ls -la $O/prof_${TAG}_fused.ncu-rep
du -sh $O
