"""`_gs`-compatible module: the 20 names of the reference's pybind extension
(gs/src/bindings.cpp:5-67), same positional signatures, in-place outputs, `None` return -- backed by
libgs3d_b200.so through the C ABI (include/gs3d_b200.h).

Hot-path bindings are real, and so is the RGB start/end pair behind
GaussianRenderer.render_aabb_culling (8f rank 2).  The remaining legacy bindings (RGB v0/v1/v2 kernels
on the CSR `offset` layout -- whose blank-tile fill is commented out in the reference,
aabb_culling.h:178-182 --, per-tile O(N*tiles) counting, image-level sort) raise NotImplementedError
naming the supported replacement; nothing falls back to a CPU implementation.

`install()` registers this module as `sys.modules["_gs"]` so the reference's
`try: import _gs as _backend` (gs/renderer.py:20-23, gs/sh_renderer.py:24-27) picks it up.
"""
import sys

import torch

from . import ops

__all__ = [
    "culling_gaussian_bsphere", "count_num_gaussians_each_tile",
    "count_num_gaussians_each_tile_bcircle", "prepare_image_sort", "image_sort",
    "tile_based_vol_rendering", "tile_based_vol_rendering_backward", "debug_check_tiledepth",
    "tile_culling_aabb", "tile_based_vol_rendering_v1", "tile_based_vol_rendering_v2",
    "tile_culling_aabb_start_end", "tile_based_vol_rendering_start_end",
    "tile_based_vol_rendering_backward_start_end", "tile_based_vol_rendering_sh",
    "tile_based_vol_rendering_backward_sh", "tile_based_vol_rendering_backward_sh_v1",
    "tile_based_vol_rendering_backward_sh_warp_reduce", "tile_based_vol_rendering_sh_with_bg",
    "tile_based_vol_rendering_backward_sh_with_bg",
]

# exact-decision mode of the compositing kernels (see include/gs3d_b200.h); module-level switch so
# the reference-signature functions stay signature-compatible.
EXACT_DECISIONS = True


def culling_gaussian_bsphere(mean, qvec, svec, normal, pts, mask, thresh):
    """bindings.cpp:6 -> render.cu:15-43."""
    ops.culling_gaussian_bsphere(mean, qvec, svec, normal, pts, mask, thresh)


def tile_culling_aabb_start_end(aabb_topleft, aabb_bottomright, gaussian_ids, start, end, depth,
                                n_tiles_h, n_tiles_w):
    """bindings.cpp:27 -> render.cu:380-397.  depth is [N,1] or [N]."""
    ops.tile_culling_aabb_start_end(aabb_topleft, aabb_bottomright, gaussian_ids, start, end, depth,
                                    n_tiles_h, n_tiles_w)


def _records(mean, cov, alpha):
    return ops.pack_records(mean, cov.reshape(-1, 4) if cov.dim() == 3 else cov, alpha)


def tile_based_vol_rendering_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft,
                                c2w, tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H,
                                W, C, thresh):
    """bindings.cpp:42 -> render.cu:483-544."""
    ops._chk(sh_coeffs, "sh_coeffs", torch.float32)
    ops.composite_sh_forward(_records(mean, cov, alpha), sh_coeffs.view(-1, 3, C * C), start, end,
                             gaussian_ids, out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
                             pixel_size_x, pixel_size_y, H, W, C, thresh, exact=EXACT_DECISIONS)


def tile_based_vol_rendering_backward_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                         grad_mean, grad_cov, grad_sh_coeffs, grad_alpha, grad_out,
                                         topleft, c2w, tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                                         pixel_size_y, H, W, C, thresh):
    """bindings.cpp:44 -> render.cu:546-624."""
    ops._chk(sh_coeffs, "sh_coeffs", torch.float32)
    ops._chk(grad_sh_coeffs, "grad_sh_coeffs", torch.float32)
    ops.composite_sh_backward(_records(mean, cov, alpha), sh_coeffs.view(-1, 3, C * C), start, end,
                              gaussian_ids, out, grad_out, grad_mean, grad_cov,
                              grad_sh_coeffs.view(-1, 3, C * C), grad_alpha, topleft, c2w, tile_size,
                              n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh,
                              exact=EXACT_DECISIONS)


def tile_based_vol_rendering_sh_with_bg(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                        topleft, c2w, tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                                        pixel_size_y, H, W, C, thresh, bg_rgb):
    """bindings.cpp:57 (vol_render_bg.h:12-100; uncompilable at reference HEAD, semantics from
    source: empty tile -> bg, else out*T + bg*(1-T) when T > thresh)."""
    ops._chk(sh_coeffs, "sh_coeffs", torch.float32)
    ops._chk(bg_rgb, "bg_rgb", torch.float32)
    ops.composite_sh_forward(_records(mean, cov, alpha), sh_coeffs.view(-1, 3, C * C), start, end,
                             gaussian_ids, out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
                             pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb=bg_rgb,
                             exact=EXACT_DECISIONS)


def tile_based_vol_rendering_backward_sh_with_bg(mean, cov, sh_coeffs, alpha, start, end,
                                                 gaussian_ids, out, grad_mean, grad_cov,
                                                 grad_sh_coeffs, grad_alpha, grad_out, topleft, c2w,
                                                 tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                                                 pixel_size_y, H, W, C, thresh, bg_rgb):
    """bindings.cpp:60 (vol_render_bg.h:121-234): the reference reuses the non-bg math on the
    blended image, so does this."""
    tile_based_vol_rendering_backward_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                         grad_mean, grad_cov, grad_sh_coeffs, grad_alpha, grad_out,
                                         topleft, c2w, tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                                         pixel_size_y, H, W, C, thresh)


def tile_based_vol_rendering_start_end(mean, cov, color, alpha, start, end, gaussian_ids, out, topleft,
                                       tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W,
                                       thresh):
    """bindings.cpp:29 -> render.cu:399-423 (vol_render.h:716-798)."""
    ops._chk(color, "color", torch.float32)
    ops.composite_rgb_forward(_records(mean, cov, alpha), color, start, end, gaussian_ids, out, topleft,
                              tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh,
                              exact=EXACT_DECISIONS)


def tile_based_vol_rendering_backward_start_end(mean, cov, color, alpha, start, end, gaussian_ids, out,
                                                grad_mean, grad_cov, grad_color, grad_alpha, grad_out,
                                                topleft, tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                                                pixel_size_y, H, W, thresh):
    """bindings.cpp:31 -> render.cu:425-481 (vol_render.h:800-923)."""
    ops._chk(color, "color", torch.float32)
    ops.composite_rgb_backward(_records(mean, cov, alpha), color, start, end, gaussian_ids, out, grad_out,
                               grad_mean, grad_cov, grad_color, grad_alpha, topleft, tile_size, n_tiles_h,
                               n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh, exact=EXACT_DECISIONS)


def debug_check_tiledepth(offset_cpu, tiledepth_cpu):
    """gs/src/debug.h:1-32: host-side sortedness check of (tile<<32|depth) keys per tile range."""
    off = offset_cpu.cpu().tolist()
    keys = tiledepth_cpu.cpu()
    for t in range(len(off) - 1):
        seg = keys[off[t]:off[t + 1]]
        if seg.numel() > 1 and not bool((seg[1:] >= seg[:-1]).all()):
            raise RuntimeError(f"tile {t}: keys not sorted")


def _legacy(name, instead):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"_gs.{name}: legacy binding outside the SH hot path (SURVEY.md 8b 'adjacent/legacy'); "
            f"use {instead}.")
    fn.__name__ = name
    fn.__doc__ = f"Legacy reference binding (not on the hot path). Use {instead}."
    return fn


def _ranges_from_offset(offset):
    """CSR offsets [n_tiles + 1] of the deprecated bindings -> (start, end) with -1 for empty tiles."""
    ops._chk(offset, "offset", torch.int32)
    start, end = offset[:-1].clone(), offset[1:].clone()
    empty = start >= end
    start[empty] = -1
    end[empty] = -1
    return start, end


def tile_based_vol_rendering(mean, cov, color, alpha, offset, gaussian_ids, out, topleft, tile_size, n_tiles_h,
                             n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh):
    """bindings.cpp:15 -> render.cu:178-217 (vol_render.h:354-412): the CSR-`offset` form of the RGB compositing
    (tile t owns gaussian_ids[offset[t]:offset[t+1]]); same per-pixel arithmetic as the start/end form
    (vol_render_one_batch, vol_render.h:252-352), so it runs on the same kernel."""
    start, end = _ranges_from_offset(offset)
    tile_based_vol_rendering_start_end(mean, cov, color, alpha, start, end, gaussian_ids, out, topleft, tile_size,
                                       n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh)


# bindings.cpp:22,25: the v1 / v2 kernels differ from v0 in their shared-memory staging only (vol_render.h:414-640)
tile_based_vol_rendering_v1 = tile_based_vol_rendering
tile_based_vol_rendering_v2 = tile_based_vol_rendering


def tile_based_vol_rendering_backward(mean, cov, color, alpha, offset, gaussian_ids, out, grad_mean, grad_cov,
                                      grad_color, grad_alpha, grad_out, topleft, tile_size, n_tiles_h, n_tiles_w,
                                      pixel_size_x, pixel_size_y, H, W, thresh):
    """bindings.cpp:17 -> render.cu:303-360: backward of the CSR-`offset` form."""
    start, end = _ranges_from_offset(offset)
    tile_based_vol_rendering_backward_start_end(mean, cov, color, alpha, start, end, gaussian_ids, out, grad_mean,
                                                grad_cov, grad_color, grad_alpha, grad_out, topleft, tile_size,
                                                n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh)


def tile_culling_aabb(aabb_topleft, aabb_bottomright, gaussian_ids, offset, depth, n_tiles_h, n_tiles_w):
    """bindings.cpp:21 -> render.cu:362-378 (aabb_culling.h:105-190): binning into the CSR-`offset` form.  The
    reference leaves the offsets of blank tiles unwritten (aabb_culling.h:178-182, the reason the form is
    deprecated); here offset is the proper exclusive scan of the per-tile counts, offset[n_tiles] = N_with_dub."""
    ops._chk(offset, "offset", torch.int32)
    n_tiles = int(n_tiles_h) * int(n_tiles_w)
    if offset.numel() != n_tiles + 1:
        raise RuntimeError("offset must have n_tiles_h * n_tiles_w + 1 elements")
    start = torch.empty(n_tiles, dtype=torch.int32, device=depth.device)
    end = torch.empty_like(start)
    ops.tile_culling_aabb_start_end(aabb_topleft, aabb_bottomright, gaussian_ids, start, end, depth,
                                    n_tiles_h, n_tiles_w)
    counts = (end - start).clamp_(min=0)  # (-1) - (-1) = 0 for blank tiles
    offset[0] = 0
    torch.cumsum(counts, 0, out=offset[1:])


count_num_gaussians_each_tile = _legacy("count_num_gaussians_each_tile", "gs.culling.tile_culling_aabb_count")
count_num_gaussians_each_tile_bcircle = _legacy("count_num_gaussians_each_tile_bcircle", "gs.culling.tile_culling_aabb_count")
prepare_image_sort = _legacy("prepare_image_sort", "tile_culling_aabb_start_end")
image_sort = _legacy("image_sort", "tile_culling_aabb_start_end")
# the two experimental backward variants compute the same gradients as the main one
tile_based_vol_rendering_backward_sh_v1 = tile_based_vol_rendering_backward_sh
tile_based_vol_rendering_backward_sh_warp_reduce = tile_based_vol_rendering_backward_sh


def install():
    sys.modules["_gs"] = sys.modules[__name__]
    return sys.modules[__name__]
