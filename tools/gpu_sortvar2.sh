#!/bin/bash
# per variant: random-pair sort 1M / 16M (back to back) and the in-engine sort phase on refdisk_1m / plummer_16m
one() {  # $1 = label, BH_LIB from env
  a=$(timeout 600 python tools/sort_bench.py 2>/dev/null | grep "onesweep_ms_back_to_back" | sed 's/.*: //' | tr -d ',\n' | sed 's/  */ /g')
  b=$(timeout 300 python tools/force_time.py refdisk_1m 2>&1 | grep split | sed "s/.*'sort': \([0-9.]*\).*/\1/")
  c=$(timeout 300 python tools/force_time.py plummer_16m 2>&1 | grep split | sed "s/.*'sort': \([0-9.]*\).*/\1/")
  echo "$1 random1M/16M: $a | engine 1M: $b 16M: $c"
}
one default
for so in nbody-barnes-hut-cuda_b200/variants/libbh_sort*.so; do BH_LIB=$PWD/$so one $(basename $so); done
