#!/bin/bash
# 2-GPU pass: the sharding test on hardware + the bench under torchrun (weak headline + strong-scaling section)
cd "$(dirname "$0")/../.."
TAG=r2e
python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi_$TAG.log 2>&1
tail -3 gpurun_out/pytest_multi_$TAG.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) \
    > gpurun_out/bench_2gpu_$TAG.json 2> gpurun_out/bench_2gpu_$TAG.err
tail -c 400 gpurun_out/bench_2gpu_$TAG.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_2gpu_$TAG.json').read().strip().split('\n')[-1])
print('value', d['value'], 'ms', d['ms_per_step'], d['ms_per_step_by_rank'])
print(json.dumps(d.get('strong'), indent=1))
PY
