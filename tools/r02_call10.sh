#!/bin/bash
mkdir -p gpurun_out
{
echo "== table resolution, batched bench (fast, c2)"
for t in 13,12 16,12 16,16 18,16 20,16 16,20; do PTB_RCT=$t timeout 120 python bench.py --steps 320 --warmup 16 --profile --precision fast --no-gate | sed "s/^/rct $t: /"; done
echo "== table resolution, batched bench (exact, c2)"
for t in 13,12 18,16; do PTB_RCT=$t timeout 120 python bench.py --steps 160 --warmup 16 --profile --precision exact | sed "s/^/rct $t: /"; done
echo "== c4 fast"
for t in 13,12 18,16; do PTB_RCT=$t timeout 120 python bench.py --config c4 --steps 96 --warmup 16 --profile --precision fast --no-gate | sed "s/^/rct $t: /"; done
} > gpurun_out/r02_call10.log 2>&1
cat gpurun_out/r02_call10.log
