#!/bin/bash
# a SECOND 20 000-iteration run of the unmodified reference (same config, same seed): how much its own test PSNR moves from run to run
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; TAG=${TAG:-r02p}
sed -e "s|^base_exp_dir: .*|base_exp_dir: /tmp/f2b_train_ref2|" oracle/ref_config_ngp_fox.yaml > /tmp/f2b_ref2.yaml
rm -rf /tmp/f2b_train_ref2
start=$(date +%s)
timeout 1200 oracle/_ref/ref_driver --train /tmp/f2b_ref2.yaml > $O/${TAG}_train_ref.log 2>&1; echo "--- reference trainer rc=$? wall_s=$(( $(date +%s) - start ))"
grep -E "Iter: +(5000|10000|15000|20000) |^[0-9]+: |Mean psnr|Nan" $O/${TAG}_train_ref.log | cut -c1-200; cat /tmp/f2b_train_ref2/train_info.txt
