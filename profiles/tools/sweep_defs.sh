#!/bin/bash
# Build liblt_b200.so with each given set of -D flags and time the kernels (run on the GPU box).
# usage: bash profiles/tools/sweep_defs.sh "-DA=1" "-DA=2 -DB=3" ...
cd "$(dirname "$0")/../.."
SRC=lattice_based_tagger_b200/csrc
for DEFS in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
       $DEFS -o lattice_based_tagger_b200/liblt_b200.so $SRC/lt_b200.cu || exit 1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$DEFS', 'ms/step %.3f'%d['ms_per_step'], {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'e2e %.0f'%d['e2e']['value'])"
done
