#!/bin/bash
# What the driver runs at round end, in one call: GPU suite, smoke, the default bench line and the reference arm.
mkdir -p gpurun_out
{
echo skip suite
echo skip smoke
echo "== bench (defaults)"; S=$(date +%s); timeout 600 python bench.py > gpurun_out/r02_head_bench_c2_n1.json 2> gpurun_out/r02_head_bench_c2_n1.err; echo "wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_head_bench_c2_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','clocks')},'gate',d['precision_gate']['passed'],'exact',round(d['exact']['value']),'roof',{k:d['roofline'][k] for k in ('achieved','frac','kernel_ms','traffic')},'e2e',round(d['e2e']['value']),d['e2e']['last_frame_on_host_equals_device_image'],'cpu',d['cpu_baseline'],'gl',d['gl_proxy']['fast'])
PY
echo "== reference arm"; S=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/r02_head_reference_arm.json 2> gpurun_out/r02_head_reference_arm.err; echo "wall $(( $(date +%s) - S )) s"; cut -c1-400 gpurun_out/r02_head_reference_arm.json
} > gpurun_out/r02_final2.log 2>&1
cat gpurun_out/r02_final2.log | cut -c1-1200
