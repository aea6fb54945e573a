#!/bin/bash
mkdir -p gpurun_out
{
echo "== GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15
echo "== bench c2 (profile mode)"; timeout 300 python bench.py --steps 320 --warmup 16 --profile
} > gpurun_out/r02_call5.log 2>&1
tail -n 30 gpurun_out/r02_call5.log
