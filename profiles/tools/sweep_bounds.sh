#!/bin/bash
# Build liblt_b200.so with several __launch_bounds__ min-blocks settings and time the two kernels.
# usage (on the GPU box): bash profiles/tools/sweep_bounds.sh "4 5 6 8" "4 5 6 8"
cd "$(dirname "$0")/../.."
SRC=lattice_based_tagger_b200/csrc
for L in $1; do for B in $2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
       -DLT_LAT_MINB=$L -DLT_BEAM_MINB=$B -o lattice_based_tagger_b200/liblt_b200.so $SRC/lt_b200.cu || exit 1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('LAT_MINB=$L BEAM_MINB=$B', 'ms/step %.3f'%d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'e2e %.0f'%d['e2e']['value'])"
done; done
