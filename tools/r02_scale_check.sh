#!/bin/bash
# N-GPU validation of HEAD the way the driver launches it: [multi-GPU tests at N=2] + bench.py under torchrun with default steps
N=${1:-2}
mkdir -p gpurun_out
{
if [ "$N" = "2" ]; then echo "== multi-GPU tests"; timeout 240 python -m pytest tests/test_multigpu_gpu.py -q -x 2>&1 | tail -n 8; fi
echo "== bench c2 N=$N"
t0=$(date +%s)
PTB_BENCH_LOG=1 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_head_bench_c2_n$N.json 2> gpurun_out/r02_head_bench_c2_n$N.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
tail -c 1200 gpurun_out/r02_head_bench_c2_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_head_bench_c2_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','steps','gpu_launches','clocks')})
    print('gate',d['precision_gate']['passed']); print('exact',d['exact'])
    print('e2e',d['e2e']['value'],d['e2e']['last_frame_on_host_equals_device_image'],d['e2e'].get('pcie_d2h_gbps')); print('xchg',d['exchange_check'])
except Exception as e: print('no line', e)
PY
} > gpurun_out/r02_scale_check_n$N.log 2>&1
tail -n 40 gpurun_out/r02_scale_check_n$N.log
