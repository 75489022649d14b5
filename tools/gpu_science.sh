#!/bin/bash
mkdir -p gpurun_out
python tools/theta_sweep.py plummer_1m > gpurun_out/theta_sweep_plummer_1m.json 2> gpurun_out/theta.err; echo "theta rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/theta_sweep_plummer_1m.json'))
for r in d['rows']: print(r)"; tail -2 gpurun_out/theta.err
python tools/energy_drift.py > gpurun_out/energy_drift.json 2> gpurun_out/energy.err; echo "energy rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/energy_drift.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, 'max drift', v['max_abs_rel_drift'], 'final', v['final_rel_drift'])"; tail -2 gpurun_out/energy.err
