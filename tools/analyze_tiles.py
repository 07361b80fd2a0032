#!/usr/bin/env python
"""Workload statistics of the compositing stage on a synthetic config (GPU): list lengths, where each
pixel / warp (8x4 px) / tile stops, and how much resident-warp time is spent waiting for the slowest
warp of the tile.   python tools/analyze_tiles.py [cfg2]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gaussian_splatting_3d_b200 import ops  # noqa: E402
from gaussian_splatting_3d_b200 import synthetic as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = "cuda:0"
cam = S.make_camera(name)
sc = S.make_scene(name, seed=0)
C = sc["C"]
r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=C))
with torch.no_grad():
    r(sc["c2w"].to(dev), cam)
st = r._state if hasattr(r, "_state") else None
# re-run the pieces by hand to get n_contrib
k1 = ops.project_cull_fused(r.mean, r.qvec, r.svec_before_activation, r.alpha_before_activation, 1, 1,
                            sc["c2w"].to(dev), cam, 1.0, False, 6.0, 16)
H, W = cam.h, cam.w
nth, ntw = (H + 15) // 16, (W + 15) // 16
ids = torch.empty(k1["n_dub"], dtype=torch.int32, device=dev)
start = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
end = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
ops.tile_culling_aabb_start_end(k1["tl"], k1["br"], ids, start, end, k1["depth"], nth, ntw, check_count=False)
out = torch.zeros(H * W * 3, device=dev)
fT = torch.zeros(H * W, device=dev)
nc = torch.zeros(H * W, dtype=torch.int32, device=dev)
topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], device=dev)
ops.composite_sh_forward(k1["records"], r.sh_coeffs, start, end, ids, out, topleft, sc["c2w"].to(dev), 16, nth, ntw,
                         1 / cam.fx, 1 / cam.fy, H, W, C, 1e-4, final_T=fT, n_contrib=nc)
torch.cuda.synchronize()
n_this = (end - start).clamp_min(0).float()
print(f"{name}: tiles {nth * ntw}, n_dub {k1['n_dub']}, list length mean {n_this.mean():.0f} max {n_this.max():.0f}")
pad = torch.zeros(nth * 16, ntw * 16, dtype=torch.int32, device=dev)
pad[:H, :W] = nc.view(H, W)
sat = torch.zeros(nth * 16, ntw * 16, device=dev)
sat[:H, :W] = (fT.view(H, W) < 1e-4).float()
t = pad.view(nth, 16, ntw, 16).permute(0, 2, 1, 3)            # [nth, ntw, 16, 16]
tile_end = t.reshape(nth, ntw, -1).max(-1).values.float()
# warps own 8x4 blocks: lx = 8*(w&1)+..., ly = 4*(w>>1)+...
wv = t.reshape(nth, ntw, 4, 4, 2, 8).permute(0, 1, 2, 4, 3, 5).reshape(nth, ntw, 8, 32)
warp_end = wv.max(-1).values.float()                            # last contributing index per warp
print(f"pixels saturated (T<1e-4): {sat[:H, :W].mean() * 100:.1f}%  mean last-contributor index per pixel "
      f"{nc.float().mean():.0f}")
print(f"tile stop index: mean {tile_end.mean():.0f} (of list {n_this.mean():.0f}); warp stop index mean "
      f"{warp_end.mean():.0f}")
print(f"sum(warp_end)/sum(8*tile_end) = {warp_end.sum() / (8 * tile_end.sum()):.3f}  "
      f"(1.0 = all warps of a tile stop together)")
b64 = torch.ceil(tile_end / 64) * 64
print(f"staged (64-batches to tile stop): {b64.sum():.0f} Gaussians = {b64.sum() / k1['n_dub'] * 100:.1f}% of n_dub")
q = torch.tensor([0.1, 0.5, 0.9, 0.99], device=dev)
print("tile_end quantiles", torch.quantile(tile_end.flatten(), q).tolist())
print("warp_end/tile_end quantiles", torch.quantile((warp_end / tile_end.clamp_min(1).unsqueeze(-1)).flatten(), q).tolist())
