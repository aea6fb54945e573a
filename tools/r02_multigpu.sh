#!/bin/bash
# Round-2 multi-GPU call (after tools/r02_first_call.sh validated batching on one GPU):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 300 -- 'bash tools/r02_multigpu.sh 8'
# Each bench run is ~15 s on the box; the call is charged N x its duration.
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 240 --warmup 10 > gpurun_out/r02_bench_n${N}_${name}.json 2> gpurun_out/r02_bench_n${N}_${name}.err
  python - "$name" gpurun_out/r02_bench_n${N}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:>14}: value {d['value']:9.0f}  back_to_back {d['back_to_back']['value']:9.0f}  e2e {d['e2e']['value']:9.0f} ({d['e2e'].get('format')}, verified {d['e2e'].get('last_frame_on_host_equals_device_image')})")
except Exception as exc:
    print(f"{sys.argv[1]:>14}: no line ({exc})")
PY
}
run default PTB_BATCH=1
run batch4 PTB_BATCH=4
run batch8 PTB_BATCH=8 PTB_SLOTS=8          # a batch cannot be longer than the exchange's slot ring
run griddiv2 PTB_GRID_DIV=2 PTB_OVERLAP=4
run nccl PTB_EXCHANGE=nccl
python tools/multigpu_check.py 2>&1 | tail -n 12
