#!/bin/bash
mkdir -p gpurun_out
{
echo "== tests"; timeout 600 python -m pytest tests -q -m gpu -x -k "atmosphere or c2_full_size" 2>&1 | tail -n 12
echo skip
} > gpurun_out/r02_call8.log 2>&1
tail -n 30 gpurun_out/r02_call8.log
