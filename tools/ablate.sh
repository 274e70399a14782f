#!/bin/bash
# builds ablation variants of the library (timing experiments only; results of the variants are NOT valid)
set -e
cd "$(dirname "$0")/../elprep_b200/csrc"
mkdir -p ../lib/exp
SRCS="api.cu sort.cu markdup.cu coordsort.cu bqsr_gather.cu bqsr_apply.cu bqsr_finalize.cu"
FL="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -shared"
build() { name=$1; shift; nvcc $FL "$@" -o ../lib/exp/lib_$name.so $SRCS -lcudart & }
rm -f ../lib/exp/*.so
for v in "$@"; do
  case $v in
    base) build base ;;
    *) build $v $(echo $v | tr '+' '\n' | sed 's/^/-DEXP_/' | tr '\n' ' ') ;;
  esac
done
wait
ls ../lib/exp
