#!/bin/bash
# Run on the GPU box through gpurun:  gpurun -- 'bash scripts/gpu_check.sh TAG [full]'
# 1. GPU parity tests  2. full bench line (+ the reference arm)  3. ncu launch list + ncu --set full of the two hot kernels.
TAG=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
python bench.py ${BENCH_ARGS} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2> gpurun_out/bench_reference_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
    k=d['kernels']; print('fps',round(d['value']),'step_frac',round(k['step_frac_of_peak'],3),'embed_ms',round(k['embed_ms'],3),round(k['embed_GBs']),'extract_ms',round(k['extract_ms'],3),round(k['extract_GBs']),'vote_ms',round(k['vote_ms'],4))
    print('e2e',{x:round(d[x]['value']) for x in d if x.startswith('e2e')},'cpu',d.get('cpu_baseline',{}).get('value'),'parity',d.get('parity',{}).get('raw_bit_mismatches_off_boundary'),'clocks',d['clocks'])
    r=json.loads(open('gpurun_out/bench_reference_${TAG}.json').read().strip().splitlines()[-1]); print('reference arm',r['value'],r['cpu_baseline']['frames'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench_${TAG}.err').read()[-2000:])
PY
if [ "$2" == "full" ]; then
CMD="python bench.py --steps 2 --warmup 1 --frames 3000 --no-e2e --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu1_${TAG}.log 2>&1
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwtsvd -s 2 -c 2 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu2_${TAG}.log 2>&1
tail -2 gpurun_out/ncu2_${TAG}.log
fi
