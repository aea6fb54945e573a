#!/bin/bash
# Round-2 opener: ONE gpurun call (~2-3 min on the box) that validates everything written after round 1's GPU minutes ran
# out and runs the queued A/B experiments.  Build the variants HERE first (they travel with the snapshot):
#   python tools/build_variants.py t128:PTB_THREADS=128,PTB_MIN_BLOCKS=8 t64:PTB_THREADS=64,PTB_MIN_BLOCKS=16
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/r02_first_call.sh'
mkdir -p gpurun_out
O=gpurun_out/r02_first
{
echo "== 1. gated GPU test of the nvcc-compiled reference shader + the config-3 golden (never run on a GPU in round 1)"
PTB_TEST_GLSL_CUDA=1 timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -k "nvcc or config3 or reference_golden" 2>&1 | tail -n 5
echo "== 1b. randomised CUDA-vs-oracle dispatches (gated in round 1)"
PTB_TEST_FUZZ=1 timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -k "randomised" 2>&1 | tail -n 6
echo "== 2. GL-compute proxy from the reference's source (exact + fast builds), 1080p"
timeout 120 python tools/gl_proxy_probe.py --frames 50
echo "== 3. frame-tail experiments: CTA size x frames in flight (us/frame; 1920x135 = the 8-GPU share)"
for v in "" variants/t128.so variants/t64.so; do
  if [ -z "$v" ] || [ -f "$v" ]; then echo "-- PTB_LIB=${v:-default}"; PTB_LIB=$v PTB_OVERLAPS=1,2,3,4 timeout 150 python tools/small_probe.py; fi
done
echo "-- default library, half / third grids (co-resident frames)"
PTB_GRID_DIV=2 PTB_OVERLAPS=2,3,4 timeout 150 python tools/small_probe.py
PTB_GRID_DIV=3 PTB_OVERLAPS=3,4 timeout 150 python tools/small_probe.py
echo "-- frame batching (one launch per 4 / 8 / 16 frames)"
for b in 4 8 16; do PTB_BATCH=$b PTB_OVERLAPS=2 timeout 150 python tools/small_probe.py; done
echo "-- gated batch tests"
PTB_TEST_BATCH=1 timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "batch" 2>&1 | tail -n 5
echo "== 4. full GPU suite"
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 4
echo "== 5. bench"
timeout 200 python bench.py --steps 200 --warmup 10
} > $O.log 2>&1
tail -n 60 $O.log
