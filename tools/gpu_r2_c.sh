#!/bin/bash
# round-2 visit: parity suite, bench both arms, ncu evidence (launch lists 1M/16M, full-set force kernel 1M and 16M)
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_ours.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
WORKLOADS="refdisk_1m plummer_16m" bash tools/gpu_launchlist2.sh
for W in refdisk_1m plummer_16m; do
  ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 1 -c 1 -f -o gpurun_out/prof_force_$W \
      python tools/profile_step.py --workload $W --steps 2 > gpurun_out/ncu_full_$W.log 2>&1
  echo "full $W rc=$?"
done
ls -la gpurun_out | head -40
