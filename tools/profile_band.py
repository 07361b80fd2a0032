#!/usr/bin/env python
"""Stage timing of one rank's share of the tile-sharded render (cfg 5), emulated on ONE GPU:
python tools/profile_band.py [world=8] [rank=3]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gaussian_splatting_3d_b200 import ops, parallel as P  # noqa: E402
from gaussian_splatting_3d_b200 import synthetic as S  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
cam = S.make_camera("cfg5")
sc = S.make_scene("cfg5", seed=0)
r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=4))
r.eval()
c2w = sc["c2w"].to(dev)
tile = 16
H, W = cam.h, cam.w
nth = (H + 15) // 16


def stage(name, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return out


for it in range(4):
    acc = {}
    k1 = stage("K1 (replicated)", lambda: ops.project_cull_fused(
        r.mean.data, r.qvec.data, r.svec_before_activation.data, r.alpha_before_activation.data, r._svec_code,
        r._alpha_code, c2w, cam, r.frustum_culling_radius, r.skip_frustum_culling, r.tile_culling_radius, tile,
        cnt=None, want_records=True, want_activated=False), acc)
    rc = stage("row_duplicate_counts", lambda: P.row_duplicate_counts(k1["tl"], k1["br"], nth), acc)
    bands = stage("balanced_bands (host)", lambda: P.balanced_bands(rc, world), acc)
    r0, r1 = bands[rank]
    img = stage("render_band", lambda: P.render_band(r, c2w, cam, r0, r1, k1=k1)[0], acc)
    if it == 3:
        print(f"world {world} rank {rank}: band rows {r0}..{r1} of {nth}; n_dub total {k1['n_dub']}")
        for k, v in acc.items():
            print(f"  {k:28s} {v:7.3f} ms")
        # inside render_band
        acc2 = {}
        tl, br, dpt, index, nb = stage("clip_rects_to_rows", lambda: ops.clip_rects_to_rows(
            k1["tl"], k1["br"], k1["depth"], r0, r1), acc2)
        ids = torch.empty(nb, dtype=torch.int32, device=dev)
        st = torch.empty(nth * ((W + 15) // 16), dtype=torch.int32, device=dev)
        en = torch.empty_like(st)
        stage("K2 binning (band)", lambda: ops.tile_culling_aabb_start_end(tl, br, ids, st, en, dpt, nth,
                                                                           (W + 15) // 16, check_count=False), acc2)
        ids = stage("ids -> original indices", lambda: torch.index_select(index, 0, ids), acc2)
        print(f"  Gaussians in the band {tl.size(0)} of {r.N}")
        out = stage("zeros(image)", lambda: torch.zeros(H * W * 3, device=dev), acc2)
        topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], device=dev)
        stage("K3 composite (band)", lambda: ops.composite_sh_forward(
            k1["records"], r.sh_coeffs.data, st, en, ids, out, topleft, c2w, tile, nth, (W + 15) // 16, 1 / cam.fx,
            1 / cam.fy, H, W, 4, r.T_thresh), acc2)
        print(f"  band duplicates {nb}")
        for k, v in acc2.items():
            print(f"    {k:26s} {v:7.3f} ms")
