#!/bin/bash
cd "$(dirname "$0")/../.."
TAG=r2f
bash profiles/tools/run_variants.sh $TAG c2
bash profiles/tools/run_variants.sh $TAG c3 --sentences 20000
