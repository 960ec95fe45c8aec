#!/bin/bash
# build_variants/build.sh NAME [extra nvcc flags...]  ->  build_variants/NAME.so  (kernel experiments; *.so is git-ignored)
set -e
cd "$(dirname "$0")/../rl_on_manifold_b200/csrc"
NAME=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --prec-div=false --prec-sqrt=false --ftz=true \
     -Xcompiler -fPIC -shared "$@" -o ../../build_variants/$NAME.so atacom_kernels.cu 2>&1 | grep -E "error|spill|Used" || true
