#!/bin/bash
mkdir -p gpurun_out
python -c "
import nbody_barnes_hut_cuda_b200 as bh
print('fp32 scalar FMA probe TFLOP/s', bh.probe_fp32_tflops(0)); print('fp32x2 packed FMA probe TFLOP/s', bh.probe_fp32x2_tflops(0)); print('hbm copy GB/s', bh.probe_hbm_gbs(0))" 2>&1 | tee gpurun_out/probes.txt
python tests/golden/make_golden.py reference 2>&1 | tail -2
python tools/sort_bench.py > gpurun_out/sort_bench.json 2>gpurun_out/sort_bench.err; cat gpurun_out/sort_bench.json; tail -3 gpurun_out/sort_bench.err
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -x 2>&1 | tail -5
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ours.json')); print('ms/step',d['ms_per_step'],'value',d['value'],'phases',d['phase_ms'],'frac',d['roofline']['frac'],'e2e ms',d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_ours.err
