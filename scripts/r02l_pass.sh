#!/bin/bash
# last round-2 GPU pass: the new bit-identity test of the background march build, the sampler tests, a 2000-iteration trainer run of
# the final C++ host (regression check of the drop-in after the last host-side changes)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02l}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "sampler" > $O/${TAG}_tests.log 2>&1
echo "--- sampler tests rc=$?"; tail -n 3 $O/${TAG}_tests.log | cut -c1-200
sed -e "s|^base_exp_dir: .*|base_exp_dir: /tmp/f2b_train_final|" -e "s|^  end_iter: .*|  end_iter: 2000|" oracle/ref_config_ngp_fox.yaml > /tmp/f2b_final.yaml
rm -rf /tmp/f2b_train_final
F2B_SHIM_PROFILE=1 timeout 600 oracle/_ref/ref_driver_b200 --train /tmp/f2b_final.yaml > $O/${TAG}_train_b200.log 2>&1; echo "--- b200 trainer rc=$?"
grep -E "Iter: +2000 |Mean psnr|f2b_shim_profile|Nan" $O/${TAG}_train_b200.log | cut -c1-220; cat /tmp/f2b_train_final/train_info.txt
