#!/bin/bash
# round-2 visit: parity suite, phase timing on three workloads, sort vs CUB
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for wl in ${WORKLOADS:-refdisk_1m plummer_1m plummer_16m}; do timeout 600 python tools/force_time.py $wl 2>&1 | grep -v "^buckets" | tail -1; done | tee gpurun_out/phases_b.txt
timeout 600 python tools/sort_bench.py > gpurun_out/sort_vs_cub.json 2> gpurun_out/sort_bench.err; cat gpurun_out/sort_vs_cub.json; tail -3 gpurun_out/sort_bench.err
for so in nbody-barnes-hut-cuda_b200/variants/libbh_sort*.so; do BH_LIB=$PWD/$so timeout 600 python tools/sort_bench.py 2>/dev/null | grep -v "cub_GB\|alg"; done
