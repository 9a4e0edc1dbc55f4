#!/bin/bash
# gpurun --gpus 8 -- 'bash scripts/scaling_run.sh TAG'   -> gpurun_out/scale_TAG_{1,2,4,8}.json
TAG=${1:-dev}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_1.json 2> gpurun_out/scale_${TAG}_1.err
python bench.py --gpus 1 --size 4k --steps 20 --warmup 3 --no-e2e > gpurun_out/scale4k_${TAG}_1.json 2> gpurun_out/scale4k_${TAG}_1.err
for N in 2 4 8; do
  E2E=$([ $N == 8 ] && echo "" || echo "--no-e2e")          # the host-buffer figure at 1 and 8 GPUs only (box time)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $E2E > gpurun_out/scale_${TAG}_$N.json 2> gpurun_out/scale_${TAG}_$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 \
  bench.py --gpus 8 --size 4k --steps 20 --warmup 3 --no-e2e > gpurun_out/scale4k_${TAG}_8.json 2> gpurun_out/scale4k_${TAG}_8.err
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/scale_${TAG}_{n}.json').read().strip().splitlines()[-1])
        base=base or d['value']
        print(n, round(d['value']), round(d['value']/base,2), round(d['ms_per_step'],3), round(d['kernels']['vote_and_combine_ms'],3), round(d.get('e2e', {}).get('value', 0)))
    except Exception as e:
        print(n, 'failed', e)
PY
python - <<PY
import json
base=None
for n in (1,8):
    try:
        d=json.loads(open(f'gpurun_out/scale4k_${TAG}_{n}.json').read().strip().splitlines()[-1])
        base=base or d['value']
        print('4k', n, round(d['value']), round(d['value']/base,2), round(d['ms_per_step'],3), round(d['kernels']['step_frac_of_peak'],3))
    except Exception as e:
        print('4k', n, 'failed', e)
PY
