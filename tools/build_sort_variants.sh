#!/bin/bash
# Tuning helper: libbh variants that differ only in bh_sort.cu compile-time knobs (see tools/build_variants.sh).
#   tools/build_sort_variants.sh name1:"-DBH_SORT_ITEMS=24 -DBH_SORT_MIN_CTAS=1" ...
set -e
cd "$(dirname "$0")/.."
PKG=nbody-barnes-hut-cuda_b200
make -C $PKG/csrc -j8 >/dev/null
mkdir -p $PKG/variants $PKG/build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
OTHERS=$(ls $PKG/build/*.o | grep -v "bh_sort.o\|bh_ic_host")
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 -I include $flags -Xptxas -v -c -o $PKG/build/var/sort_$name.o $PKG/csrc/bh_sort.cu 2>&1 \
      | grep -A2 "onesweep_kernelILb0ELb0" | grep "Used\|spill" | tr '\n' ' ' | sed "s/^/$name: /"; echo
    nvcc $ARCH -shared -o $PKG/variants/libbh_sort_$name.so $OTHERS $PKG/build/var/sort_$name.o -ldl ) &
done
wait
