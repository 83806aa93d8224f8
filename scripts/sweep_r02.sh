#!/bin/bash
# round-2 measurement sweep (one gpurun call): backward ray chunks x resident scatter CTAs per SM, with the march pipelined
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep_r02.jsonl
for chunks in 1 2 4; do
  for ctas in 8 2 1; do
    F2B_BWD_CHUNKS=$chunks F2B_SCATTER_CTAS=$ctas timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu-baseline 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'bwd_chunks': $chunks, 'scatter_ctas': $ctas, 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'hash_bwd_ms': d['kernels'].get('f2b_hash_bwd',{}).get('ms_per_step'), 'mlp_bwd2_ms': d['kernels'].get('f2b_mlp_bwd2',{}).get('ms_per_step')}))" >> gpurun_out/sweep_r02.jsonl
  done
done
cat gpurun_out/sweep_r02.jsonl
