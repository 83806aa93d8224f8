#!/bin/bash
# sixth round-2 GPU pass (2 GPUs): hash scatter level rotation (tests + bench), then weak / strong scaling lines at N = 2
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02f}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -q -p no:cacheprovider -k "hash or headline or train_step" > $O/${TAG}_tests.log 2>&1
echo "--- hash / headline tests rc=$?"; tail -n 4 $O/${TAG}_tests.log | cut -c1-250
TAG=$TAG bash scripts/r02_multi_gpu.sh 2
du -sh $O
