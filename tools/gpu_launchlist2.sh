#!/bin/bash
# ncu launch list (per-kernel durations, DRAM bytes) of 2 direct-launch steps for the given workloads
mkdir -p gpurun_out
for W in ${WORKLOADS:-refdisk_1m}; do
python tools/profile_step.py --workload $W --steps 2 > gpurun_out/plain_$W.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv \
    python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_launches_$W.log 2>&1
echo "$W launch list rc=$?"
done
