#!/bin/bash
# one --set full capture of every non-force kernel of a step (second step: bodies already sorted)
mkdir -p gpurun_out
W=${WORKLOAD:-refdisk_1m}
python tools/profile_step.py --workload $W --steps 2 > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|link_kernel|com_cells_kernel|com_scan_kernel|scan_pairs_kernel|init_cells_kernel|integrate_kernel|keys_kernel|histogram_kernel' -f -o gpurun_out/prof_small_$W \
    python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_small.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_small.log; ls -la gpurun_out/prof_small_$W.ncu-rep
