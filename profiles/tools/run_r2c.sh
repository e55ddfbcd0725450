#!/bin/bash
cd "$(dirname "$0")/../.."
TAG=r2c
bash profiles/tools/run_variants.sh $TAG c2
bash profiles/tools/run_variants.sh $TAG c3 --sentences 20000
python -m pytest tests -q -m gpu -x -k "random_cases or golden or config_samples or lookup_modes" > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
