#!/bin/bash
mkdir -p gpurun_out
{
echo "== GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 15
echo "== rct probe"; timeout 300 python tools/rct_probe.py
} > gpurun_out/r02_call2.log 2>&1
tail -n 40 gpurun_out/r02_call2.log
