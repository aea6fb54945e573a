#!/bin/bash
mkdir -p gpurun_out
{
echo "== large-scene tests"; timeout 600 python -m pytest tests -q -m gpu -x  2>&1 | tail -n 12
echo "== C3 probes"
timeout 120 python tools/c3_probe.py
PTB_PRECISION=fast timeout 120 python tools/c3_probe.py
PTB_LARGE=bvh timeout 120 python tools/c3_probe.py
} > gpurun_out/r02_call6.log 2>&1
tail -n 30 gpurun_out/r02_call6.log
