#!/bin/bash
# Run on the GPU box through gpurun:  gpurun -- 'bash scripts/gpu_check.sh TAG [full]'
# 1. GPU parity tests  2. full bench line  3. ncu launch list + ncu --set full of the two hot kernels.
TAG=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
    k=d['kernels']; print('fps',round(d['value']),'step_frac',round(k['step_frac_of_peak'],3),'embed_ms',round(k['embed_ms'],3),round(k['embed_GBs']),'extract_ms',round(k['extract_ms'],3),round(k['extract_GBs']),'vote_ms',round(k['vote_and_combine_ms'],4))
    print('e2e',d.get('e2e',{}).get('value'),'cpu',d.get('cpu_baseline',{}).get('value'),'acc',d['bit_accuracy'],'clocks',d['clocks'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench_${TAG}.err').read()[-2000:])
PY
if [ "$2" == "full" ]; then
CMD="python bench.py --steps 2 --warmup 1 --frames 300 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu1_${TAG}.log 2>&1
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwtsvd -s 2 -c 2 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu2_${TAG}.log 2>&1
tail -2 gpurun_out/ncu2_${TAG}.log
fi
