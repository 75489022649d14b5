#!/bin/bash
# quick GPU check: parity suite (-x) then a short bench without the CPU baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -6
timeout 600 python bench.py --steps ${STEPS:-30} --warmup 3 --no-cpu-baseline --no-scale-ref ${BENCH_ARGS} > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ours.json')); print('ms/step',d['ms_per_step'],'value',d['value'],'phases',d['phase_ms'],'frac',d['roofline']['frac'],'int/body',d['interactions_per_body'],'e2e ms',d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_ours.err
