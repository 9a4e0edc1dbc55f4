#!/bin/bash
# Tuning aid.  HERE (build container):  bash scripts/tune_variants.sh build TAG "-DB200WM_X=1 ..." [TAG2 "..."]...
#   builds one libb200wm.so per flag set into gpurun_variants/ (travels with the snapshot, git-ignored).
# On the GPU box:  bash scripts/tune_variants.sh run   -> one short bench line per variant.
set -e
cd "$(dirname "$0")/.."
if [ "$1" == "build" ]; then
  shift
  mkdir -p gpurun_variants
  while [ $# -gt 1 ]; do
    B200WM_NVCC_EXTRA="$2" python video-fingerprinting_b200/build.py --force > /dev/null
    cp video-fingerprinting_b200/lib/libb200wm.so gpurun_variants/libb200wm_$1.so
    echo "built $1: $2"
    shift 2
  done
  python video-fingerprinting_b200/build.py --force > /dev/null     # leave the default build in place
else
  cp video-fingerprinting_b200/lib/libb200wm.so /tmp/libb200wm_default.so
  for v in gpurun_variants/libb200wm_*.so; do
    cp $v video-fingerprinting_b200/lib/libb200wm.so
    python bench.py ${VARIANT_BENCH_ARGS} --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$v', 'fps', round(d['value']), 'embed_ms', round(k['embed_ms'],3), 'extract_ms', round(k['extract_ms'],3), 'acc', d['bit_accuracy']['frames_exact'])"
  done
  cp /tmp/libb200wm_default.so video-fingerprinting_b200/lib/libb200wm.so
fi
