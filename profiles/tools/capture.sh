#!/bin/bash
# One round of measurement artefacts, run on the GPU box:  bash profiles/tools/capture.sh <tag>
#   gpurun_out/bench_c2_<tag>.json         default bench line (C2 headline + api + other_configs, CPU baseline and parity)
#   gpurun_out/bench_reference_<tag>.json  the reference arm (the reference's own Tagger.tag on all host cores)
#   gpurun_out/launches_<tag>.csv          ncu launch list (gpu__time_duration.sum) of a short C2 bench run
#   gpurun_out/prof_<tag>.ncu-rep          ncu --set full of the last beam / lattice launches of that run (C2)
#   gpurun_out/prof_c3_<tag>.ncu-rep       the same for the C3 sample (1 M-entry dictionary, beam 10)
# Every ncu pass starts only after the plain run of the same command exited 0.
cd "$(dirname "$0")/../.."
TAG=${1:-run}
SHORT="--steps 2 --warmup 3 --no-cpu-baseline --no-api --other-configs none"
set -x
if [ -z "$NCU_ONLY" ]; then
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err || exit 1
fi
python bench.py $SHORT > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|lattice_kernel" -s 10 -c 4 -f \
    -o gpurun_out/prof_$TAG python bench.py $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
python bench.py --config c3 --sentences 20000 $SHORT > /dev/null 2>&1 || exit 1
# (every launch of the two kernels: the warm-up of this configuration holds grow-and-rerun rounds and retry passes, the
# steady-state launches are the longest ones — ncu_summary.py picks those)
ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|lattice_kernel" -c 80 -f \
    -o gpurun_out/prof_c3_$TAG python bench.py --config c3 --sentences 20000 $SHORT > gpurun_out/ncu_full_c3_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
