#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nproc; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
PTB_BENCH_LOG=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 320 --warmup 16 > gpurun_out/r02_bench_c2_n${N}_numa.json 2> gpurun_out/r02_bench_c2_n${N}_numa.err
python - gpurun_out/r02_bench_c2_n${N}_numa.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'exact', d['exact'] and d['exact']['value'], 'e2e', d['e2e']['value'], d['e2e']['last_frame_on_host_equals_device_image'], d['e2e'].get('host_frame_numa'), 'xchg', d['exchange_check'] and d['exchange_check']['exchange_equals_single_gpu'])
except Exception as e:
    print('no line', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
