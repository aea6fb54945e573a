#!/bin/bash
# Profiling call (one GPU): launch list of the bench command + ncu full captures of the headline kernels.
mkdir -p gpurun_out
{
echo "== launch list (bench, fast, batched)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 120 --csv --log-file gpurun_out/r02_launches_bench_fast.csv python bench.py --steps 96 --warmup 16 --profile --precision fast --no-gate 2>&1 | tail -n 2
echo "== ncu full: C2 fast, one batched launch (16 frames)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c2_fast_batch16 python bench.py --steps 48 --warmup 16 --profile --precision fast --no-gate 2>&1 | tail -n 3
echo "== ncu full: C2 exact, one batched launch (16 frames)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c2_exact_batch16 python bench.py --steps 48 --warmup 16 --profile --precision exact 2>&1 | tail -n 3
echo "== ncu full: C3 grid (exact)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c3_grid python tools/c3_probe.py 2>&1 | tail -n 3
echo "== ncu full: batch blend kernel"
timeout 300 ncu --set full --clock-control none -k regex:blend_batch -s 2 -c 1 -f -o gpurun_out/r02_blend_batch python bench.py --steps 48 --warmup 16 --profile --precision fast --no-gate 2>&1 | tail -n 3
} > gpurun_out/r02_profile.log 2>&1
tail -n 30 gpurun_out/r02_profile.log
