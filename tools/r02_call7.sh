#!/bin/bash
mkdir -p gpurun_out
{
echo "== grid density sweep (C3)"
for dens in 1 2 4 8; do PTB_GRID_DENSITY=$dens timeout 120 python tools/c3_probe.py; done
for dens in 2 4; do PTB_PRECISION=fast PTB_GRID_DENSITY=$dens timeout 120 python tools/c3_probe.py; done
echo "== full GPU suite"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 6
echo "== bench c4 N=1"; timeout 300 python bench.py --config c4 --steps 96 --warmup 16 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -c 300 gpurun_out/r02_bench_c4_n1.err
echo "== bench c3 N=1"; timeout 300 python bench.py --config c3 --steps 64 --warmup 16 > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3_n1.err; tail -c 300 gpurun_out/r02_bench_c3_n1.err
python - <<'PY'
import json
for n in ('c4','c3'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{n}_n1.json').read().strip().splitlines()[-1])
        print(n,{k:d[k] for k in ('value','ms_per_step')},'gate',d['precision_gate']['per_channel_mse_fast_vs_exact'],d['precision_gate']['passed'],'exact',d['exact'] and d['exact']['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'])
    except Exception as e: print(n,'no line',e)
PY
} > gpurun_out/r02_call7.log 2>&1
tail -n 40 gpurun_out/r02_call7.log
