#!/bin/bash
# A/B of the compositing-backward variants (build/ablate/libgs3d_exp.so, -DGS3D_BWD_EXPERIMENTS)
mkdir -p gpurun_out; rm -f /tmp/grad_ref.pt
for v in ${VARIANTS:-0 1 2 3 4 5 6}; do
  GS3D_GRAD_REF=/tmp/grad_ref.pt GS3D_LIB=build/ablate/libgs3d_exp.so GS3D_BWD_VARIANT=$v timeout 300 python tools/bench_composite.py cfg2 10 2>&1 | tail -1
done | tee gpurun_out/bwd_variants.jsonl
