#!/bin/bash
# build_variant.sh <name> [nvcc -D flags...]  ->  lattice_based_tagger_b200/variants/<name>.so  (experiments only)
cd "$(dirname "$0")/../.."
NAME=$1; shift
mkdir -p lattice_based_tagger_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC "$@" \
    -o lattice_based_tagger_b200/variants/$NAME.so lattice_based_tagger_b200/csrc/lt_b200.cu
