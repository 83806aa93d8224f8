#!/bin/bash
# eighth round-2 GPU pass: the round's final build — whole suite, bench line, ncu launch list + --set full, where a trainer
# iteration goes (C++ host stopwatch through 4000 iterations of the unmodified trainer)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02h}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $O/${TAG}_pytest_gpu.log 2>&1
echo "--- pytest -m gpu (all): rc=$?"; tail -n 4 $O/${TAG}_pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "--- smoke rc=$?"; tail -n 1 $O/${TAG}_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "--- bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    k = {n: round(v["ms_per_step"], 3) for n, v in list(d["kernels"].items())[:14]}
    print(json.dumps({"ms_per_step": d["ms_per_step"], "value": d["value"], "e2e_ms": d["e2e"]["ms_per_step"], "attempts": [a["rejected"] for a in d["timing_attempts"]],
                      "e2e_attempts": [a["rejected"] for a in d["e2e"]["timing_attempts"]], "roofline": d["roofline"], "kernels": k, "clocks": d["clocks"],
                      "forward_only": d.get("forward_only"), "reference_gpu": d.get("reference_gpu"), "cpp_host": d.get("cpp_host"), "cpu_baseline": d.get("cpu_baseline"),
                      "launches": d["gpu_launches"]}))
except Exception as e:
    print("parse failed", e); print(open("$O/${TAG}_bench.err").read()[-1500:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "--- bench --impl reference rc=$?"; tail -c 400 $O/${TAG}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_bench_ncu.log 2>&1
echo "--- ncu launch list rc=$?"; wc -l $O/${TAG}_launches.csv
timeout 1200 ncu --set full --import-source on --clock-control none -k "regex:march16|hash_bwd|field_fwd|mlp_fwd_tc|mlp_bwd|shader_prep_bwd|composite_bwd|composite_fwd|compact|mark_visit|early_stop" -c 17 -f -o $O/prof_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu > $O/${TAG}_ncu_full.log 2>&1
echo "--- ncu --set full rc=$?"; ls -la $O/prof_$TAG.ncu-rep
sed -e "s|^base_exp_dir: .*|base_exp_dir: /tmp/f2b_prof_train|" -e "s|^  end_iter: .*|  end_iter: 4000|" oracle/ref_config_ngp_fox.yaml > /tmp/f2b_prof.yaml
rm -rf /tmp/f2b_prof_train
( nvidia-smi --query-gpu=utilization.gpu --format=csv,noheader,nounits -lms 250 > $O/${TAG}_util_b200.txt & echo $! > /tmp/smi.pid )
start=$(date +%s.%N); F2B_SHIM_PROFILE=1 timeout 900 oracle/_ref/ref_driver_b200 --train /tmp/f2b_prof.yaml > $O/${TAG}_train_prof.log 2>&1; end=$(date +%s.%N)
kill $(cat /tmp/smi.pid) 2>/dev/null
echo "--- shim profile (4000 iterations, wall $(python -c "print(round($end-$start,1))") s, train_info $(cat /tmp/f2b_prof_train/train_info.txt 2>/dev/null))"; grep f2b_shim_profile $O/${TAG}_train_prof.log
python -c "
v=[int(x) for x in open('$O/${TAG}_util_b200.txt').read().split() if x.strip().isdigit()]
print('gpu utilisation samples', len(v), 'mean', round(sum(v)/max(len(v),1),1))"
du -sh $O
