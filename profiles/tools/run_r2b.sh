#!/bin/bash
cd "$(dirname "$0")/../.."
TAG=r2b
( time python -m pytest tests -q -m gpu ) > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
bash profiles/tools/run_variants.sh $TAG c2
bash profiles/tools/run_variants.sh $TAG c3 --sentences 20000
