#!/bin/bash
# What the round-end driver does on a fresh box: GPU tests, smoke, both bench arms with default flags.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s)-T0 )) s)"; tail -2 gpurun_out/pytest_gpu.log
T0=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; echo "smoke ($(( $(date +%s)-T0 )) s)"
T0=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/BENCH_ref.json 2> gpurun_out/BENCH_ref.err; echo "ref rc=$? ($(( $(date +%s)-T0 )) s)"; wc -l gpurun_out/BENCH_ref.json
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/BENCH_ours.json 2> gpurun_out/BENCH_ours.err; echo "ours rc=$? ($(( $(date +%s)-T0 )) s)"; wc -l gpurun_out/BENCH_ours.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/BENCH_ref.json')); o=json.load(open('gpurun_out/BENCH_ours.json'))
print('reference', round(r['value']/1e6,2),'M body-steps/s', round(r['ms_per_step'],2),'ms; cpu', round(r['cpu_baseline']['value']/1e6,2),'M')
print('ours value', round(o['value']/1e6,1),'M', round(o['ms_per_step'],3),'ms; e2e', round(o['e2e']['value']/1e6,1),'M', round(o['e2e']['ms_per_step'],3),'ms; roofline', o['roofline']['achieved'], o['roofline']['peak'], o['roofline']['frac'], 'traffic', o['roofline']['traffic'])
print('ratio value', o['value']/r['value'], 'ratio e2e', o['e2e']['value']/r['value'])
print('phases', o['phase_ms']); print('scale_ref', o.get('scale_ref_1gpu')); print('clocks', o['clocks'], 'launches', o['gpu_launches'])
PY
