#!/bin/bash
for f in build/ablate/libgs3d_*.so; do GS3D_LIB=$f python tools/bench_composite.py cfg2 10 2>&1 | tail -1; done
