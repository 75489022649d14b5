#!/bin/bash
# one --set full capture of a named kernel on a workload (tools/profile_step.py must have passed plain first)
mkdir -p gpurun_out
W=${WORKLOAD:-refdisk_1m}; K=${KERNEL:-force_kernel}; S=${SKIP:-4}
python tools/profile_step.py --workload $W --steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_${K}_$W \
    python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
