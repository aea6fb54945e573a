#!/bin/bash
# Final single-GPU pass: full suite, smoke, bench lines for c2 / c3 / c4, launch list + ncu captures of the headline kernels.
mkdir -p gpurun_out
{
echo "== GPU suite"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 6
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
for cfg in c2 c3 c4; do
  echo "== bench $cfg"; steps=320; [ $cfg = c3 ] && steps=64; [ $cfg = c4 ] && steps=96
  timeout 400 python bench.py --config $cfg --steps $steps --warmup 16 > gpurun_out/r02_final_bench_${cfg}_n1.json 2> gpurun_out/r02_final_bench_${cfg}_n1.err; tail -c 300 gpurun_out/r02_final_bench_${cfg}_n1.err
done
python - <<'PY'
import json
for n in ('c2','c3','c4'):
    try:
        d=json.loads(open(f'gpurun_out/r02_final_bench_{n}_n1.json').read().strip().splitlines()[-1])
        print(n,{k:d[k] for k in ('value','ms_per_step','gpu_launches')},'gate',d['precision_gate']['passed'],[f"{x:.2e}" for x in d['precision_gate']['per_channel_mse_fast_vs_exact']],'exact',d['exact'] and round(d['exact']['value']),'roof',{k:d['roofline'][k] for k in ('achieved','frac','kernel_ms','kernel_ms_per_launch')},'e2e',round(d['e2e']['value']),d['e2e']['last_frame_on_host_equals_device_image'],'cpu',round(d['cpu_baseline']['value'],2),'gl',d.get('gl_proxy',{}).get('fast',{}).get('msamples_per_s'))
    except Exception as e: print(n,'no line',e)
PY
echo "== launch list (bench, fast, batched)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 120 --csv --log-file gpurun_out/r02_launches_bench_fast.csv python bench.py --steps 96 --warmup 16 --profile --precision fast --no-gate 2>&1 | tail -n 1
echo "== ncu full: C2 fast / exact, one batched launch (16 frames)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c2_fast_batch16 python bench.py --steps 48 --warmup 16 --profile --precision fast --no-gate 2>&1 | tail -n 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c2_exact_batch16 python bench.py --steps 48 --warmup 16 --profile --precision exact 2>&1 | tail -n 1
} > gpurun_out/r02_final1.log 2>&1
tail -n 30 gpurun_out/r02_final1.log
