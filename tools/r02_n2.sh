#!/bin/bash
# 2-GPU validation: multi-GPU tests + a short bench (gate, exchange check, scatter e2e)
mkdir -p gpurun_out
{
echo "== multi-GPU tests"; timeout 200 python -m pytest tests/test_multigpu_gpu.py -q -x 2>&1 | tail -n 8
echo "== bench c2 N=2"
PTB_BENCH_LOG=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 160 --warmup 16 > gpurun_out/r02_bench_c2_n2_a.json 2> gpurun_out/r02_bench_c2_n2_a.err
tail -c 1500 gpurun_out/r02_bench_c2_n2_a.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_c2_n2_a.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
    print('gate',d['precision_gate']); print('exact',d['exact']); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','kernel_ms','kernel_ms_per_launch','frames_per_launch')})
    print('e2e',d['e2e']['value'],d['e2e']['last_frame_on_host_equals_device_image']); print('xchg',d['exchange_check'])
except Exception as e: print('no line', e)
PY
} > gpurun_out/r02_n2.log 2>&1
tail -n 40 gpurun_out/r02_n2.log
