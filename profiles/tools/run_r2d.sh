#!/bin/bash
cd "$(dirname "$0")/../.."
TAG=r2d
export LT_DEBUG=1
bash profiles/tools/run_variants.sh $TAG c2
bash profiles/tools/run_variants.sh $TAG c3 --sentences 20000
grep -h "beam kernel\|lattice kernel" gpurun_out/var_${TAG}_c2_*.err | sort | uniq -c | head -20
