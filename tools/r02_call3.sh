#!/bin/bash
mkdir -p gpurun_out
{
echo "== GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x -s -k "fast or classification" 2>&1 | tail -n 15
echo "== rct probe exact"; timeout 300 python tools/rct_probe.py 13,12 13,16
echo "== rct probe fast"; PTB_PRECISION=fast timeout 300 python tools/rct_probe.py 13,12 13,16 18,16
echo "== ncu fast"; PTB_PRECISION=fast bash tools/ncu_c2.sh r02_c2_fast_rct
echo "== ncu exact"; bash tools/ncu_c2.sh r02_c2_exact_rct
} > gpurun_out/r02_call3.log 2>&1
tail -n 40 gpurun_out/r02_call3.log
