#!/bin/bash
# ncu full capture (with source correlation) of the C2 megakernel: one launch after warm-up.
mkdir -p gpurun_out
NAME=${1:-r02_c2_mega}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 5 -c 1 -f -o gpurun_out/$NAME python tools/perf_probe_c2.py 2>&1 | tail -n 5
