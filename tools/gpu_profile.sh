#!/bin/bash
# ncu evidence: launch list (per-kernel share of a step) + one full capture of the force kernel.
mkdir -p gpurun_out
W=${WORKLOAD:-refdisk_1m}
python tools/profile_step.py --workload $W --steps 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv \
    python tools/profile_step.py --workload $W --steps 3 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-force_kernel} -s 1 -c 1 -f -o gpurun_out/prof_${KERNEL:-force}_$W \
    python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
