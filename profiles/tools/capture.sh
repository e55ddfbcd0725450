#!/bin/bash
# One round of measurement artefacts, run on the GPU box:  bash profiles/tools/capture.sh <tag>
#   gpurun_out/bench_c2_<tag>.json         default bench line (C2, with the CPU baseline and parity check)
#   gpurun_out/bench_reference_<tag>.json  the reference arm (CPU oracle port, all host cores)
#   gpurun_out/launches_<tag>.csv          ncu launch list (gpu__time_duration.sum) of a short bench run
#   gpurun_out/prof_<tag>.ncu-rep          ncu --set full of the last beam / lattice launches of that run
# Every ncu pass starts only after the plain run of the same command exited 0.
cd "$(dirname "$0")/../.."
TAG=${1:-run}
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|lattice_kernel" -s 8 -c 6 -f \
    -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
