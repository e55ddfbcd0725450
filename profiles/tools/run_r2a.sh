#!/bin/bash
# round-2 first GPU pass: GPU tests, default bench (with other_configs), reference arm, ncu launch list + full set (C2)
cd "$(dirname "$0")/../.."
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
tail -c 600 gpurun_out/bench_c2_$TAG.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --other-configs none > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --other-configs none > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|lattice_kernel" -s 10 -c 4 -f \
    -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api --other-configs none > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
