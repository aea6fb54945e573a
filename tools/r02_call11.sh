#!/bin/bash
mkdir -p gpurun_out
{
for dens in 2 3 6; do PTB_GRID_DENSITY=$dens timeout 120 python tools/c3_probe.py; done
echo "== bench c3"; timeout 300 python bench.py --config c3 --steps 64 --warmup 16 > gpurun_out/r02_final_bench_c3_n1.json 2> gpurun_out/r02_final_bench_c3_n1.err; tail -c 300 gpurun_out/r02_final_bench_c3_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final_bench_c3_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')},'gate',d['precision_gate']['passed'],d['precision_gate']['per_channel_mse_fast_vs_exact'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'])
PY
echo "== ncu c3 grid"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c3_grid python tools/c3_probe.py 2>&1 | tail -n 2
} > gpurun_out/r02_call11.log 2>&1
cat gpurun_out/r02_call11.log | cut -c1-300
