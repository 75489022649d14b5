#!/bin/bash
mkdir -p gpurun_out
W=${WORKLOAD:-plummer_16m}
python tools/profile_step.py --workload $W --steps 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv \
    python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain.log
