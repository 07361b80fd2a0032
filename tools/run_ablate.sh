#!/bin/bash
# times every build/ablate/libgs3d_*.so (tools/ablate.sh) with tools/bench_composite.py; gradients are compared with the first build's
rm -f /tmp/grad_ref.pt
for f in build/ablate/libgs3d_*.so; do GS3D_GRAD_REF=/tmp/grad_ref.pt GS3D_LIB=$f timeout 300 python tools/bench_composite.py ${CFG:-cfg2} 10 2>&1 | tail -1; done
