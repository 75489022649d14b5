#!/bin/bash
# full-set capture of the tree-build kernels only (second step), both workloads
mkdir -p gpurun_out
for W in refdisk_1m plummer_16m; do
ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|link_kernel|scan_pairs_kernel|init_cells_kernel|scan_tiles_kernel' --launch-skip 5 -f -o gpurun_out/prof_build_$W \
    python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_build.log 2>&1
echo "rc=$?"; python tools/ncu_brief.py gpurun_out/prof_build_$W.ncu-rep
done
