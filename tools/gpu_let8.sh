#!/bin/bash
# 8-GPU LET visit: small two-disc system first (validates the 8-rank path), then configs[4] (256M bodies).
mkdir -p gpurun_out
N=${NGPU:-8}
run() {  # workload steps
  timeout ${3:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps $2 --warmup 3 --workload $1 --mode let ${BENCH_EXTRA} > gpurun_out/let_$1_n$N.json 2> gpurun_out/let_$1_n$N.err
  rc=$?; echo "$1 rc=$rc"; cat gpurun_out/let_$1_n$N.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/let_$1_n$N.err | tail -8
  return $rc
}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
{ [ "${SKIP_SMALL:-0}" = "1" ] || run twodisk_16m 5 300; } && run ${BIG:-twodisk_256m} ${BIGSTEPS:-10} 700
