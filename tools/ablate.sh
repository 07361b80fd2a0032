#!/bin/bash
# Experimental builds of the C ABI with -DGS3D_ABLATE=<bits> (C = 4 kernels only) into build/ablate/.
# usage: tools/ablate.sh 0 1 2 4 7 ...   (then GS3D_LIB=build/ablate/libgs3d_ab<N>.so python tools/bench_composite.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ablate
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -cudart shared \
    -DGS3D_ONLY_C4 -DGS3D_ABLATE=$n $EXTRA -o build/ablate/libgs3d_ab$n.so \
    gaussian_splatting_3d_b200/csrc/project.cu gaussian_splatting_3d_b200/csrc/binning.cu gaussian_splatting_3d_b200/csrc/composite.cu &
done
wait
ls -la build/ablate
