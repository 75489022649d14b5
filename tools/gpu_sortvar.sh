#!/bin/bash
timeout 600 python tools/sort_bench.py 2>/dev/null | grep -v "cub_GB\|alg"
for so in nbody-barnes-hut-cuda_b200/variants/libbh_sort*.so; do BH_LIB=$PWD/$so timeout 600 python tools/sort_bench.py 2>/dev/null | grep -v "cub_GB\|alg\|cub_ms\|l2_flushed\|^{\|^}\|^ }" | tr -d '\n'; echo; done
