#!/bin/bash
mkdir -p gpurun_out
{
echo "== GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15
echo "== rct probe exact"; timeout 300 python tools/rct_probe.py 13,12
echo "== rct probe fast"; PTB_PRECISION=fast timeout 300 python tools/rct_probe.py 13,12 18,16
echo "== bench c2"; timeout 600 python bench.py > gpurun_out/r02_bench_c2_n1_a.json 2> gpurun_out/r02_bench_c2_n1_a.err; tail -c 600 gpurun_out/r02_bench_c2_n1_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_n1_a.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('gate',d['precision_gate']); print('exact',d['exact']); print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','kernel_ms','kernel_ms_per_launch','frames_per_launch')})
print('e2e',d['e2e']['value'],d['e2e']['last_frame_on_host_equals_device_image'],d['e2e_other_formats']); print('cpu',d['cpu_baseline']); print('gl',d.get('gl_proxy'))
PY
} > gpurun_out/r02_call4.log 2>&1
tail -n 40 gpurun_out/r02_call4.log
