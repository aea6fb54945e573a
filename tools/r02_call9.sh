#!/bin/bash
mkdir -p gpurun_out
{
echo "== GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 8
echo "== A/B completion queue (in-place single-frame launches)"
for d in 0 1; do echo "PTB_DEFER=$d"; PTB_DEFER=$d timeout 120 python tools/rct_probe.py 13,12 | tail -n 1; PTB_DEFER=$d PTB_PRECISION=fast timeout 120 python tools/rct_probe.py 13,12 18,16 | tail -n 2; done
echo "== bench profile mode (batched), fast"
for d in 0 1; do PTB_DEFER=$d timeout 120 python bench.py --steps 320 --warmup 16 --profile --precision fast --no-gate | sed "s/^/defer $d: /"; done
} > gpurun_out/r02_call9.log 2>&1
tail -n 30 gpurun_out/r02_call9.log
