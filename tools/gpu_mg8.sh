#!/bin/bash
# 8-GPU visit (charged 8x): sliced bit-identity, sliced bench (16M Plummer), LET bench (256M two-disc, configs[4])
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/check_sliced.py 2>&1 | grep "SLICED_CHECK"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
   bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 > gpurun_out/scale_plummer_16m_n$N.json 2> gpurun_out/scale_plummer_16m_n$N.err
echo "sliced rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/scale_plummer_16m_n$N.json')); print('N',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'same1',d['same_workload_1gpu']['ms_per_step'],'eff',d['efficiency_same_workload'],'phases',d['phase_ms_rank0'],'force/rank',d['force_ms_per_rank'],'allgather',d['allgather_ms'],'e2e',d['e2e']['ms_per_step'])"
if [ "${LET:-1}" = "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps ${LETSTEPS:-12} --warmup 4 --workload ${LETW:-twodisk_256m} --mode let ${LET_EXTRA} > gpurun_out/let_${LETW:-twodisk_256m}_n$N.json 2> gpurun_out/let_${LETW:-twodisk_256m}_n$N.err
echo "let rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/let_${LETW:-twodisk_256m}_n$N.json')); print('LET N',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'force share',d.get('force_share'),'forces/rank',d['forces_update_ms_per_rank'],'int/body',d['interactions_per_body'],'e2e',d['e2e'] and d['e2e']['ms_per_step'],'engine',d['engine']['key_bits'],d['engine']['let_interval'])"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/let_${LETW:-twodisk_256m}_n$N.err | tail -5
fi
