#!/bin/bash
# Tuning helper: build libbh variants that differ only in bh_force.cu compile-time knobs.
#   tools/build_variants.sh name1:"-DFORCE_ITEMS=3" name2:"-DFORCE_MIN_CTAS=5" ...
# -> nbody-barnes-hut-cuda_b200/variants/libbh_<name>.so   (select with BH_LIB=...)
set -e
cd "$(dirname "$0")/.."
PKG=nbody-barnes-hut-cuda_b200
make -C $PKG/csrc -j8 >/dev/null
mkdir -p $PKG/variants $PKG/build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
OTHERS=$(ls $PKG/build/*.o | grep -v bh_force.o)
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 $flags -Xptxas -v -c -o $PKG/build/var/force_$name.o $PKG/csrc/bh_force.cu 2>&1 \
      | grep -A2 "force_kernelILi10" | grep "Used" | sed "s/^/$name: /"
    nvcc $ARCH -shared -o $PKG/variants/libbh_$name.so $OTHERS $PKG/build/var/force_$name.o ) &
done
wait
ls -la $PKG/variants/
