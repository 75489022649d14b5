#!/bin/bash
# e2e timeline + full-set ncu of one onesweep pass and one CUB pass on the same 16M random pairs
mkdir -p gpurun_out
BH_LIB=$PWD/nbody-barnes-hut-cuda_b200/variants/libbh_trace.so timeout 300 python tools/step_host_trace.py 2>&1 | tail -24
timeout 300 python tools/sort_prof.py > gpurun_out/sort_prof_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel --launch-skip 9 --launch-count 2 -f -o gpurun_out/sort_ours_16m python tools/sort_prof.py > gpurun_out/ncu_sort_ours.log 2>&1; echo "ours rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:DeviceRadixSortOnesweep --launch-skip 5 --launch-count 2 -f -o gpurun_out/sort_cub_16m python tools/sort_prof.py > gpurun_out/ncu_sort_cub.log 2>&1; echo "cub rc=$?"
ls -la gpurun_out/*.ncu-rep
