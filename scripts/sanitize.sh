#!/bin/bash
# compute-sanitizer over scripts/sanitize_case.py: bash scripts/sanitize.sh [tag]   (run under gpurun; ~10 min)
# Writes gpurun_out/sanitize_<tag>_<tool>.log and prints each tool's summary line.
# (On this project's GPU pool compute-sanitizer is switched off - the tools exit 86; the plain run of the case and the
# canary-bounds tests in tests/test_gpu_dwtsvd.py are what was run there.)
tag=${1:-r02}
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_case.py > gpurun_out/sanitize_${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_${tag}_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_${tag}_plain.log
for tool in memcheck racecheck synccheck initcheck; do
    args=""
    [ "$tool" != memcheck ] && args="--quick"
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $args > gpurun_out/sanitize_${tag}_${tool}.log 2>&1
    echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tag}_${tool}.log | tail -1) ; $(grep -c 'sanitize case done' gpurun_out/sanitize_${tag}_${tool}.log) completed"
done
