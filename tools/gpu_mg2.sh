#!/bin/bash
# 2..N GPU visit: C++/NCCL sliced driver — bit-identity check (torchrun), the C++ front-end with --gpus, a short bench line
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/check_sliced.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
echo "== nbody_bench --gpus $N"
timeout 300 ./nbody-barnes-hut-cuda_b200/nbody_bench --n 4000000 --ic plummer --frames 10 --quiet --gpus $N 2>&1 | tail -6
echo "== nbody_bench --gpus 1"
timeout 300 ./nbody-barnes-hut-cuda_b200/nbody_bench --n 4000000 --ic plummer --frames 10 --quiet 2>&1 | tail -4
echo "== bench.py --gpus $N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
   bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_mg_n$N.json 2> gpurun_out/bench_mg_n$N.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mg_n$N.json')); print('N',d['n_gpus'],'ms/step',round(d['ms_per_step'],3),'same1',d['same_workload_1gpu']['ms_per_step'],'eff',d['efficiency_same_workload'],'phases',d['phase_ms_rank0'],'force/rank',d['force_ms_per_rank'],'allgather',d['allgather_ms'],'e2e',d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_mg_n$N.err | grep -v "^\*\|OMP_NUM"
