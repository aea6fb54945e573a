#!/bin/bash
# 8-GPU: bench c2 and c4 through torchrun; raw JSON lines kept.
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, args...
  local name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N "$@" > gpurun_out/r02_bench_${name}_n${N}.json 2> gpurun_out/r02_bench_${name}_n${N}.err
  python - gpurun_out/r02_bench_${name}_n${N}.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'exact', d['exact'] and d['exact']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['last_frame_on_host_equals_device_image'], 'xchg', d['exchange_check'] and d['exchange_check']['exchange_equals_single_gpu'], 'gate', d['precision_gate'] and d['precision_gate']['passed'], d['clocks'])
except Exception as e:
    print(sys.argv[1], 'no line', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
run c2 --steps 320 --warmup 16
run c4 --config c4 --steps 160 --warmup 16
