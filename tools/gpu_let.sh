#!/bin/bash
# LET visit: emulated-rank tests on one GPU, torchrun check against the single-GPU run, LET bench leg.
mkdir -p gpurun_out
N=${NGPU:-2}; W=${WORKLOAD:-twodisk_16m}; K=${STEPS:-5}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_let.py -x -q -p no:cacheprovider 2>&1 | tail -15
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/check_let.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -8
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps $K --warmup 3 --workload $W --mode let ${BENCH_EXTRA} > gpurun_out/let_${W}_n$N.json 2> gpurun_out/let_${W}_n$N.err
echo "rc=$?"; cat gpurun_out/let_${W}_n$N.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/let_${W}_n$N.err | tail -12
