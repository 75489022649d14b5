#!/bin/bash
# 16M Plummer (BASELINE.json configs[3]) at 1/2/4/8 GPUs on one box, back to back.
mkdir -p gpurun_out
W=${WORKLOAD:-plummer_16m}; K=${STEPS:-10}
for N in ${NLIST:-1 2 4 8}; do
  if [ $N = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps $K --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/scale_${W}_n1.json 2> gpurun_out/scale_${W}_n1.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) \
       bench.py --gpus $N --steps $K --warmup 3 --workload $W > gpurun_out/scale_${W}_n$N.json 2> gpurun_out/scale_${W}_n$N.err
  fi
  echo "N=$N rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/scale_${W}_n$N.json')); print('N',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'value',round(d['value']/1e6,1),'M body-steps/s','int/s',round(d['interactions_per_s']/1e9,1),'G', d.get('phase_ms_rank0', d.get('phase_ms')), 'allgather_ms', d.get('allgather_ms'), 'e2e ms', round(d['e2e']['ms_per_step'],2))" 2>&1 | tail -1
  tail -2 gpurun_out/scale_${W}_n$N.err | grep -v "^$" | grep -iv "omp_num\|\*\*\*" | tail -2
done
