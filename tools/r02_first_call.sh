#!/bin/bash
# Round-2 opener: ONE gpurun call that validates everything written after round 1's GPU minutes ran out (the three test
# gates are gone: the suite now runs batch / fuzz / nvcc-GLSL / the fixed BVH probe) and runs the queued A/B experiments.
#   python tools/build_variants.py t128:PTB_THREADS=128,PTB_MIN_BLOCKS=8
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r02_first_call.sh'
mkdir -p gpurun_out
O=gpurun_out/r02_first
{
echo "== 0. box"; nproc; nvidia-smi -L
echo "== 1. full GPU suite (un-gated)"
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 25
echo "== 1b. remaining tests after a failure (if any)"
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 8
echo "== 2. GL-compute proxy from the reference's source (exact + fast builds), 1080p"
timeout 120 python tools/gl_proxy_probe.py --frames 50
echo "== 3. frame-tail experiments: frames in flight (us/frame; 1920x135 = the 8-GPU share)"
PTB_OVERLAPS=1,2,3 timeout 150 python tools/small_probe.py
echo "-- t128 variant"
PTB_LIB=variants/t128.so PTB_OVERLAPS=2,3 timeout 150 python tools/small_probe.py
echo "-- frame batching (one launch per 4 / 8 / 16 frames)"
for b in 4 8 16; do PTB_BATCH=$b PTB_OVERLAPS=2 timeout 150 python tools/small_probe.py; done
echo "== 4. BVH threshold A/B"
timeout 200 python tools/bvh_threshold_probe.py
echo "== 5. perf probe / C3"
timeout 200 python tools/perf_probe.py
timeout 100 python tools/c3_probe.py
echo "== 6. bench"
timeout 300 python bench.py --steps 200 --warmup 10
echo "== 7. ncu full capture of the BVH instantiation (C3)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 3 -c 1 -f -o gpurun_out/r02_c3_bvh python tools/c3_probe.py 2>&1 | tail -n 5
} > $O.log 2>&1
tail -n 120 $O.log
