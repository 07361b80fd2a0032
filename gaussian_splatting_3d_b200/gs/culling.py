"""gs/culling.py of the reference: tile AABB rects + duplicate count (culling.py:8-37)."""
import torch

from .. import ops


@torch.no_grad()
def tile_culling_aabb_count(mean, cov, tile_size, camera_info, D):
    """-> (N_with_dub: int, aabb_topleft int32 [N,2], aabb_bottomright int32 [N,2]) in tile units,
    inclusive.  One kernel + one 8-byte read-back (the reference's `.item()`), bit-exact with the
    reference's ~20 FP32 torch ops: sqrt(D*cov_xx), (m -+ e)*f + c in separate roundings,
    truncation toward zero, clamp into the image (quirk Q8), floor-divide by the tile size."""
    if not mean.is_cuda:
        raise RuntimeError("tile_culling_aabb_count: CUDA tensors required (no CPU fallback)")
    return ops.tile_culling_aabb_count(mean.contiguous(), cov.contiguous(), tile_size, camera_info, D)
