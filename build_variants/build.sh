#!/bin/bash
# build_variants/build.sh NAME [extra nvcc flags...]  ->  build_variants/NAME.so  (kernel experiments; *.so is git-ignored)
# Full ptxas log in build_variants/NAME.log; the registers / spills of the fix-up and iiwa step kernels are echoed.
cd "$(dirname "$0")/../rl_on_manifold_b200/csrc" || exit 1
NAME=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --prec-div=false --prec-sqrt=false --ftz=true \
     -Xcompiler -fPIC -shared -Xptxas -v "$@" -o ../../build_variants/$NAME.so atacom_kernels.cu > ../../build_variants/$NAME.log 2>&1
grep -E " error" ../../build_variants/$NAME.log | head
echo "built $NAME rc=$?"
