#!/usr/bin/env python
"""Kernel-only timing of K3 / K4a (compositing forward / backward) on a synthetic config with CUDA
events; used for A/B experiments (GS3D_LIB selects the build).  python tools/bench_composite.py [cfg2] [reps]"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gaussian_splatting_3d_b200 import ops  # noqa: E402
from gaussian_splatting_3d_b200 import synthetic as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda:0"
cam = S.make_camera(name)
sc = S.make_scene(name, seed=0)
C = sc["C"]
p = {k: sc[k].to(dev) for k in ("mean", "qvec", "svec_before_activation", "alpha_before_activation", "sh_coeffs")}
c2w = sc["c2w"].to(dev)
k1 = ops.project_cull_fused(p["mean"], p["qvec"], p["svec_before_activation"], p["alpha_before_activation"], 1, 1,
                            c2w, cam, 1.0, False, 6.0, 16)
H, W = cam.h, cam.w
nth, ntw = (H + 15) // 16, (W + 15) // 16
ids = torch.empty(k1["n_dub"], dtype=torch.int32, device=dev)
start = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
end = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
ops.tile_culling_aabb_start_end(k1["tl"], k1["br"], ids, start, end, k1["depth"], nth, ntw, check_count=False)
topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], device=dev)
out = torch.zeros(H * W * 3, device=dev)
gout = torch.rand(H * W * 3, device=dev) * 1e-6
N = p["mean"].size(0)
gm, gc, ga = torch.zeros(N, 2, device=dev), torch.zeros(N, 4, device=dev), torch.zeros(N, device=dev)
gsh = torch.zeros_like(p["sh_coeffs"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fwd():
    ops.composite_sh_forward(k1["records"], p["sh_coeffs"], start, end, ids, out, topleft, c2w, 16, nth, ntw,
                             1 / cam.fx, 1 / cam.fy, H, W, C, 1e-4)


def bwd():
    ops.composite_sh_backward(k1["records"], p["sh_coeffs"], start, end, ids, out, gout, gm, gc, gsh, ga, topleft,
                              c2w, 16, nth, ntw, 1 / cam.fx, 1 / cam.fy, H, W, C, 1e-4)


def timeit(fn):
    ts = []
    for i in range(reps + 2):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


res = {"lib": os.environ.get("GS3D_LIB", "default"), "variant": os.environ.get("GS3D_BWD_VARIANT", ""), "cfg": name,
       "fwd_ms": round(timeit(fwd), 4)}
gout = (2.0 * (out - torch.rand(H * W * 3, device=dev, generator=torch.Generator(dev).manual_seed(1))) / out.numel())
res["bwd_ms"] = round(timeit(bwd), 4)
# gradient fingerprint of ONE backward from zeroed buffers (A/B builds must agree to float-atomics noise)
for t in (gm, gc, gsh, ga):
    t.zero_()
bwd()
torch.cuda.synchronize()
res["img_sum"] = float(out.sum())
res["grad_norms"] = [float(t.double().norm()) for t in (gm, gc, gsh, ga)]
ref_path = os.environ.get("GS3D_GRAD_REF")
if ref_path:
    if os.path.exists(ref_path):
        ref = torch.load(ref_path, map_location=dev)
        res["grad_rel_vs_ref"] = [float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
                                  for a, b in zip((gm, gc, gsh, ga), ref)]
    else:
        torch.save([gm, gc, gsh, ga], ref_path)
print(json.dumps(res))
