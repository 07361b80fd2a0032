#!/bin/bash
# ncu --set full of the compositing-backward variants (experimental build), one launch each
mkdir -p gpurun_out
for v in ${VARIANTS:-1 7}; do
  GS3D_LIB=build/ablate/libgs3d_exp.so GS3D_BWD_VARIANT=$v timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'composite_bwd' -s 3 -c 1 -o gpurun_out/prof_bwd_v$v -f python tools/bench_composite.py cfg2 2 > gpurun_out/ncu_bwd_v$v.log 2>&1; echo "ncu v$v rc=$?"
done
ls -la gpurun_out/*.ncu-rep
