#!/bin/bash
# ncu launch list of bench.py itself (same command, short run), as B200_PROFILING.md asks.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scale-ref"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
echo "rc=$?"; tail -1 gpurun_out/bench_plain.err; wc -l gpurun_out/launches_bench.csv
