#!/bin/bash
# blend grid sweep: N GPUs (torchrun) or 1
N=${1:-2}
mkdir -p gpurun_out
for ctas in ${CTAS:-16 32 64 148}; do
  if [ "$N" = "1" ]; then
    PTB_BLEND_CTAS=$ctas timeout 300 python bench.py --steps 160 --warmup 16 --profile 2>/dev/null | tail -n 1 | sed "s/^/ctas $ctas: /"
  else
    PTB_BLEND_CTAS=$ctas timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 160 --warmup 16 --profile 2>/dev/null | tail -n 1 | sed "s/^/ctas $ctas: /"
  fi
done 2>&1 | tee gpurun_out/blend_probe_n$N.log
