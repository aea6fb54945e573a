#!/bin/bash
# compute-sanitizer memcheck + racecheck over tools/sanitizer_target.py (B200); summary into gpurun_out/r02_compute_sanitizer.md
mkdir -p gpurun_out
out=gpurun_out/r02_compute_sanitizer.md
echo "# compute-sanitizer on tools/sanitizer_target.py (round 2 kernels: table / grid / BVH folds, completion queue, batch blend, fast build; B200)" > $out
echo '```' >> $out
for tool in memcheck racecheck; do
  t0=$(date +%s)
  timeout 280 compute-sanitizer --tool $tool python tools/sanitizer_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  rc=$?
  echo "compute-sanitizer --tool $tool (rc=$rc, $(( $(date +%s) - t0 )) s): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -n 1)" >> $out
  grep -c "^done" gpurun_out/sanitizer_$tool.log | sed 's/^/  target completed: /' >> $out
done
echo '```' >> $out
cat $out
tail -n 12 gpurun_out/sanitizer_memcheck.log
