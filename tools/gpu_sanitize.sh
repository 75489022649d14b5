#!/bin/bash
# compute-sanitizer memcheck on a small workload (one tool per gpurun call, B200_PROFILING.md)
mkdir -p gpurun_out
python tools/profile_step.py --workload uniform_16k --steps 2 > gpurun_out/plain_san.log 2>&1 &&
timeout 900 compute-sanitizer --tool ${TOOL:-memcheck} --error-exitcode 9 python tools/profile_step.py --workload uniform_16k --steps 2 > gpurun_out/sanitizer_${TOOL:-memcheck}.log 2>&1
echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer_${TOOL:-memcheck}.log
