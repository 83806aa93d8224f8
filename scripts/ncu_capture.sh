#!/bin/bash
# One `ncu --set full` capture (with source correlation) of the hot kernels of one training step.
# usage: bash scripts/ncu_capture.sh <tag> [kernel-regex] [count]   -> gpurun_out/prof_<tag>.ncu-rep
TAG=${1:-r01}
RE=${2:-'march16|hash_bwd|field_fwd|mlp_fwd_tc|mlp_bwd_tc|shader_prep_bwd|composite_bwd|composite_fwd|compact|mark_visit|early_stop'}
CNT=${3:-18}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$RE" -c $CNT -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu > gpurun_out/ncu_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_$TAG.log | cut -c1-300
ls -la gpurun_out/prof_$TAG.ncu-rep
