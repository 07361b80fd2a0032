"""CameraInfo -- the input type of the hot path (reference utils/camera.py:219-323).

Same constructor, attributes and methods.  The COLMAP / blender dataset loaders of the reference
file (:326-609) are out of scope.
"""
import numpy as np
import torch
import torch.nn.functional as F


class CameraInfo:
    def __init__(self, fx, fy, cx, cy, w, h, near_plane, far_plane) -> None:
        self.fx, self.fy = fx, fy
        self.cx, self.cy = cx, cy
        self.w, self.h = w, h
        self.near_plane = near_plane
        self.far_plane = far_plane
        self._refresh()

    def _refresh(self):
        self.yfov = 2 * np.arctan(self.h / (2 * self.fy))
        self.aspect = self.w / self.h

    def downsample(self, scale):
        self.fx, self.fy = self.fx / scale, self.fy / scale
        self.cx, self.cy = self.cx / scale, self.cy / scale
        self.w //= scale
        self.h //= scale
        self._refresh()

    def upsample(self, scale):
        # reference :241-247 does not refresh yfov/aspect (both are scale invariant)
        self.fx, self.fy = self.fx * scale, self.fy * scale
        self.cx, self.cy = self.cx * scale, self.cy * scale
        self.w *= int(scale)
        self.h *= int(scale)

    def get_frustum(self, c2w):
        """Six (normal, point) planes of the view frustum, reference :249-283.  On CUDA this is one
        tiny kernel (gs3d_get_frustum) instead of ~20 ATen launches; on CPU tensors it is the
        same arithmetic in torch."""
        if c2w.is_cuda:
            from .. import ops

            return ops.get_frustum(c2w.contiguous().float(), self)
        up, right, look, t = -c2w[:, 1], c2w[:, 0], c2w[:, 2], c2w[:, 3]
        half_v = self.far_plane * np.tan(self.yfov * 0.5)
        half_h = half_v * self.aspect
        near_pt, far_pt = self.near_plane * look, self.far_plane * look
        normals = torch.stack([
            look,
            -look,
            torch.linalg.cross(far_pt - half_h * right, up),
            torch.linalg.cross(up, far_pt + half_h * right),
            torch.linalg.cross(far_pt + half_v * up, right),
            torch.linalg.cross(right, far_pt - half_v * up),
        ], dim=0)
        pts = torch.stack([near_pt + t, far_pt + t, t, t, t, t], dim=0)
        return F.normalize(normals, dim=-1), pts

    def print_camera_info(self):
        print(f"camera: fx {self.fx:.2f} fy {self.fy:.2f} cx {self.cx:.2f} cy {self.cy:.2f} "
              f"H {self.h} W {self.w} pixel_size {1 / self.fx:.4g} {1 / self.fy:.4g}")

    def camera_space_to_pixel_space(self, pts):
        """reference :290-303 (in-place scale+shift, truncation toward zero -- quirk Q8)."""
        if pts.shape[1] == 3:
            pts = pts[:, :2] / pts[:, 2:]
        assert pts.shape[1] == 2
        pts[:, 0] = pts[:, 0] * self.fx + self.cx
        pts[:, 1] = pts[:, 1] * self.fy + self.cy
        if isinstance(pts, np.ndarray):
            return pts.astype(np.int32)
        return pts.to(torch.int32)

    @classmethod
    def from_fov_camera(cls, fov, aspect, resolution, near_plane, far_plane):
        W = resolution
        H = int(resolution / aspect)
        cx, cy = W / 2, H / 2
        fx = cx / np.tan(fov / 2)
        fy = cy / np.tan(fov / 2)
        return cls(fx, fy, cx, cy, W, H, near_plane, far_plane)


def in_frustum(queries, normal, pts):
    """CPU twin of the point-in-frustum test, reference :317-323 (radius 0, strict > 0)."""
    ok = torch.ones_like(queries[..., 0], dtype=torch.bool)
    for i in range(6):
        ok &= ((queries - pts[i]) @ normal[i]) > 0.0
    return ok


def get_c2w_from_up_and_look_at(up, look_at, pos):
    """Pose helper of the reference's fixtures (gs/debug.py:23-37): numpy in, float32 [3,4] out;
    columns are x = (-up) x z, y = z x x, z = normalised view direction, then the position."""
    up = np.asarray(up, dtype=np.float32)
    z = np.asarray(look_at, dtype=np.float32) - np.asarray(pos, dtype=np.float32)
    z = z / np.linalg.norm(z)
    x = np.cross(-(up / np.linalg.norm(up)), z)
    y = np.cross(z, x)
    return np.stack([x, y, z, np.asarray(pos, dtype=np.float32)], axis=1).astype(np.float32)
