#!/bin/bash
# round-2 visit: parity suite, then force-phase timing of the tuning variants (BH_LIB) on two workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
V=nbody-barnes-hut-cuda_b200/variants
for wl in ${WORKLOADS:-refdisk_1m plummer_1m}; do
  timeout 300 python tools/force_time.py $wl 2>&1 | grep -v "^buckets" | tail -2
  for so in $V/libbh_*.so; do
    BH_LIB=$PWD/$so timeout 300 python tools/force_time.py $wl 2>&1 | grep -v "^buckets" | tail -1
  done
done 2>&1 | tee gpurun_out/variants_a.txt
for a in ${SPLITS:-}; do BH_SPLIT=$a timeout 300 python tools/force_time.py refdisk_1m 2>&1 | grep -v "^buckets" | tail -1; done | tee -a gpurun_out/variants_a.txt
