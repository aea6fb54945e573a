#!/bin/bash
# What the driver runs at round end, in one call: GPU suite, smoke, the default bench line and the reference arm.
mkdir -p gpurun_out
{
echo "== GPU suite"; S=$(date +%s); timeout 500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 6; echo "wall $(( $(date +%s) - S )) s"
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
echo "== bench (defaults)"; S=$(date +%s); timeout 400 python bench.py > gpurun_out/r02_head_bench_c2_n1.json 2> gpurun_out/r02_head_bench_c2_n1.err; echo "wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_head_bench_c2_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','clocks')},'gate',d['precision_gate']['passed'],'exact',round(d['exact']['value']),'roof',{k:d['roofline'][k] for k in ('achieved','frac','kernel_ms','traffic')},'e2e',round(d['e2e']['value']),d['e2e']['last_frame_on_host_equals_device_image'],d['e2e']['pcie_d2h_gbps'],'cpu',d['cpu_baseline'],'gl',d['gl_proxy']['fast'])
PY
echo "== reference arm"; S=$(date +%s); timeout 400 python bench.py --impl reference > gpurun_out/r02_head_reference_arm.json 2> gpurun_out/r02_head_reference_arm.err; echo "wall $(( $(date +%s) - S )) s"; cut -c1-400 gpurun_out/r02_head_reference_arm.json
} > gpurun_out/r02_final3.log 2>&1
cat gpurun_out/r02_final3.log | cut -c1-1400
