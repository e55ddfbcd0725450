#!/bin/bash
# One round of measurement artefacts, run on the GPU box:  bash profiles/tools/capture.sh <tag> [parts...]
# parts (default: all), each writing into gpurun_out/ (which travels back only while it stays below 64 MiB):
#   bench   bench_c2_<tag>.json         default bench line (C2 headline + api + other_configs, CPU baseline and parity)
#   ref     bench_reference_<tag>.json  the reference arm (the reference's own Tagger.tag on all host cores)
#   ncu_c2  launches_<tag>.csv          ncu launch list (gpu__time_duration.sum) of a short C2 bench run
#           prof_<tag>.ncu-rep          ncu --set full of the last beam / lattice launches of that run
#           <tag>_ncu_full_summary.json, <tag>_beam_by_line.txt, <tag>_lattice_by_line.txt   (made here, on the box)
#   ncu_c3  <tag>_ncu_full_summary_c3.json, <tag>_*_by_line_c3.txt of the C3 sample (1 M-entry dictionary, beam 10); the
#           report itself (hundreds of MB) is deleted after the summaries are made
# Every ncu pass starts only after the plain run of the same command exited 0.
cd "$(dirname "$0")/../.."
TAG=${1:-run}; shift
PARTS=${@:-bench ref ncu_c2 ncu_c3}
SHORT="--steps 2 --warmup 3 --no-cpu-baseline --no-api --other-configs none"
LIB=lattice_based_tagger_b200/liblt_b200.so
set -x
for PART in $PARTS; do
case $PART in
bench)
  python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err || exit 1 ;;
ref)
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err || exit 1 ;;
ncu_c2)
  python bench.py $SHORT > /dev/null 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
  # (the last launches of the run: the staged pass of the timed steps)
  ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|lattice_kernel" -s 14 -c 4 -f \
      -o gpurun_out/prof_$TAG python bench.py $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
  python profiles/tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep gpurun_out/${TAG}_ncu_full_summary.json gpurun_out/traffic_$TAG.json c2
  python profiles/tools/ncu_by_line.py gpurun_out/prof_$TAG.ncu-rep beam_kernel $LIB 60 > gpurun_out/${TAG}_beam_by_line.txt
  python profiles/tools/ncu_by_line.py gpurun_out/prof_$TAG.ncu-rep lattice_kernel $LIB 60 > gpurun_out/${TAG}_lattice_by_line.txt ;;
ncu_c3)
  python bench.py --config c3 --sentences 20000 $SHORT > /dev/null 2>&1 || exit 1
  # (every launch of the two kernels: the warm-up of this configuration holds grow-and-rerun rounds and retry passes, the
  # steady-state launches are the longest ones — ncu_summary.py picks those)
  ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|lattice_kernel" -c 80 -f \
      -o /tmp/prof_c3_$TAG python bench.py --config c3 --sentences 20000 $SHORT > gpurun_out/ncu_full_c3_$TAG.log 2>&1
  python profiles/tools/ncu_summary.py /tmp/prof_c3_$TAG.ncu-rep gpurun_out/${TAG}_ncu_full_summary_c3.json gpurun_out/traffic_$TAG.json c3
  rm -f /tmp/prof_c3_$TAG.ncu-rep ;;
esac
done
ls -la gpurun_out/*$TAG*
