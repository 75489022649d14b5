#!/bin/bash
# round-2 visit A: parity suite on the flattened traversal, then force-phase timing of the tuning variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
V=nbody-barnes-hut-cuda_b200/variants
for wl in refdisk_1m plummer_1m; do
  timeout 300 python tools/force_time.py $wl 2>&1 | grep -v "^buckets" | tail -2
  for so in $V/libbh_*.so; do
    BH_LIB=$PWD/$so timeout 300 python tools/force_time.py $wl 2>&1 | grep -v "^buckets" | tail -1
  done
done 2>&1 | tee gpurun_out/variants_a.txt
