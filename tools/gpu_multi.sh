#!/bin/bash
# Multi-GPU visit: sliced bench at N GPUs (torchrun) for one workload.
mkdir -p gpurun_out
N=${NGPU:-2}; W=${WORKLOAD:-plummer_16m}; K=${STEPS:-10}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps $K --warmup 3 --workload $W ${BENCH_EXTRA} > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err
echo "rc=$?"; cat gpurun_out/bench_${W}_n$N.json; tail -5 gpurun_out/bench_${W}_n$N.err
if [ "${ALSO1:-0}" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps $K --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/bench_${W}_n1.json 2> gpurun_out/bench_${W}_n1.err
  echo "rc1=$?"; cat gpurun_out/bench_${W}_n1.json; tail -3 gpurun_out/bench_${W}_n1.err
fi
