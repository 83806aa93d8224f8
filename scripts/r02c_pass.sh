#!/bin/bash
# third round-2 GPU pass: whole suite (reference binaries present), bench line, ncu of the new kernels, per-config lines, chunk sweep
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02c}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $O/${TAG}_pytest_gpu.log 2>&1
echo "--- pytest -m gpu (all): rc=$?"; tail -n 6 $O/${TAG}_pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "--- bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:12]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e_ms": d["e2e"]["ms_per_step"], "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]],
                      "roofline": d["roofline"], "rooflines": d["rooflines"], "kernels": k, "forward_only": d.get("forward_only"),
                      "reference_gpu": d.get("reference_gpu"), "cpu_baseline": d.get("cpu_baseline")}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench.err").read()[-1500:])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:render_fwd_fused|mlp_bwd_rc|mlp_fwd_tc" -c 8 -f -o $O/prof_${TAG}_new \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_ncu_new.log 2>&1
echo "--- ncu new kernels rc=$?"; ls -la $O/prof_${TAG}_new.ncu-rep
for C in free nerf360 big20 big22; do
  timeout 600 python bench.py --config $C --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_$C.json 2> $O/${TAG}_bench_$C.err
  echo "--- bench --config $C rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_$C.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:6]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e_ms": d["e2e"]["ms_per_step"], "measured": d["workload_measured"], "kernels": k,
                      "reference_gpu": d.get("reference_gpu"), "forward_only": (d.get("forward_only") or {}).get("ms_per_step")}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench_$C.err").read()[-800:])
PY
done
: > $O/${TAG}_sweep.jsonl
for V in "2 8" "2 2" "4 2" "4 1"; do set -- $V
  F2B_BWD_CHUNKS=$1 F2B_SCATTER_CTAS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'bwd_chunks': $1, 'scatter_ctas': $2, 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'hash_bwd_ms': d['kernels'].get('f2b_hash_bwd',{}).get('ms_per_step'), 'mlp_bwd2_ms': d['kernels'].get('f2b_mlp_bwd2',{}).get('ms_per_step')}))" >> $O/${TAG}_sweep.jsonl
done
echo "--- sweep (bwd chunks x scatter CTAs/SM)"; cat $O/${TAG}_sweep.jsonl
du -sh $O
