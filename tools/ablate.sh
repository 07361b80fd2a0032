#!/bin/bash
# Experimental builds of the C ABI (C = 4 kernels only) into build/ablate/, one per "name:flags" argument.
# usage: tools/ablate.sh base: u41:"-DGS3D_UNROLL_F=4 -DGS3D_UNROLL_B=1" ab1:-DGS3D_ABLATE=1
#        then on the GPU box: tools/run_ablate.sh   (GS3D_LIB=build/ablate/libgs3d_<name>.so python tools/bench_composite.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ablate
rm -f build/ablate/*.so
for arg in "$@"; do
  name="${arg%%:*}"; flags="${arg#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -cudart shared \
    -DGS3D_ONLY_C4 $flags -o build/ablate/libgs3d_$name.so \
    gaussian_splatting_3d_b200/csrc/project.cu gaussian_splatting_3d_b200/csrc/binning.cu gaussian_splatting_3d_b200/csrc/composite.cu gaussian_splatting_3d_b200/csrc/exchange.cu gaussian_splatting_3d_b200/csrc/bands.cu gaussian_splatting_3d_b200/csrc/train.cu gaussian_splatting_3d_b200/csrc/loss.cu &
done
wait
ls build/ablate
