"""Synthetic scene recipe of SURVEY.md 8(d) / BASELINE.md 3 (no COLMAP data offline).

All draws come from one CPU torch.Generator in a fixed order, so a (config, seed) pair names the
same scene everywhere: tests, bench.py (both arms), the oracle.  Draw order: z, u, v, qvec,
svec, alpha, SH degree 0, SH higher orders; the target image uses seed + 1.
"""
import math

import torch

from .utils.camera import CameraInfo

# name -> (N, C (= SH degree + 1), W, H, fx, fy, cx, cy)
CONFIGS = {
    "cfg1": (10_000, 1, 256, 256, 300.0, 300.0, 128.0, 128.0),
    "cfg2": (3_000_000, 4, 1297, 840, 961.22, 963.09, 648.38, 420.12),
    "cfg3": (500_000, 3, 1008, 756, 815.0, 815.0, 504.0, 378.0),
    "cfg5": (6_000_000, 4, 3840, 2160, 2846.0, 2846.0, 1920.0, 1080.0),
}

RENDER_DEFAULTS = dict(
    tile_size=16, frustum_culling_radius=1.0, tile_culling_radius=6.0, T_thresh=1e-4,
    near_plane=1.5, far_plane=1000.0,
)


def make_camera(name):
    _, _, W, H, fx, fy, cx, cy = CONFIGS[name]
    return CameraInfo(fx, fy, cx, cy, W, H, RENDER_DEFAULTS["near_plane"], RENDER_DEFAULTS["far_plane"])


def make_scene(name=None, seed=0, N=None, C=None, camera=None, max_C=None):
    """-> dict of CPU float32 tensors: mean [N,3], qvec [N,4], svec_before_activation [N,3],
    alpha_before_activation [N], sh_coeffs [N,3,max_C^2], c2w [3,4]; plus 'C' and 'camera'."""
    if name is not None:
        n0, c0 = CONFIGS[name][0], CONFIGS[name][1]
        N = n0 if N is None else N
        C = c0 if C is None else C
        camera = make_camera(name) if camera is None else camera
    max_C = C if max_C is None else max_C
    g = torch.Generator().manual_seed(seed)
    W, H, fx, fy, cx, cy = camera.w, camera.h, camera.fx, camera.fy, camera.cx, camera.cy
    z = 2.0 + 10.0 * torch.rand(N, generator=g)
    u = -0.55 + 1.1 * torch.rand(N, generator=g)
    v = -0.55 + 1.1 * torch.rand(N, generator=g)
    x = u * (W / fx) * z + ((W / 2 - cx) / fx) * z
    y = v * (H / fy) * z + ((H / 2 - cy) / fy) * z
    mean = torch.stack([x, y, z], dim=-1).float()
    qvec = torch.randn(N, 4, generator=g)
    lo, hi = math.log(0.004), math.log(0.04)
    svec_ba = lo + (hi - lo) * torch.rand(N, 3, generator=g)
    alpha_ba = 1.5 * torch.randn(N, generator=g)
    sh = torch.zeros(N, 3, max_C * max_C)
    sh[:, :, 0] = torch.randn(N, 3, generator=g) / 0.2821
    if max_C > 1:
        sh[:, :, 1:] = 0.3 * torch.randn(N, 3, max_C * max_C - 1, generator=g)
    c2w = torch.eye(3, 4)
    return dict(mean=mean, qvec=qvec, svec_before_activation=svec_ba, alpha_before_activation=alpha_ba,
                sh_coeffs=sh, c2w=c2w, C=C, camera=camera)


def make_target(camera, seed=0):
    g = torch.Generator().manual_seed(seed + 1)
    return torch.rand(camera.h, camera.w, 3, generator=g)


def ring_cameras(n=8, radius=7.0, centre=(0.0, 0.0, 7.0)):
    """cfg 4: n c2w [3,4] poses on a ring around the scene centre, looking at it (OpenCV axes:
    x right, y down, z forward)."""
    out = []
    c = torch.tensor(centre)
    for i in range(n):
        a = 2 * math.pi * i / n
        pos = c + radius * torch.tensor([math.sin(a), 0.0, -math.cos(a)])
        zf = (c - pos) / torch.linalg.norm(c - pos)
        up = torch.tensor([0.0, -1.0, 0.0])
        xr = torch.linalg.cross(zf, up)
        xr = xr / torch.linalg.norm(xr)
        yd = torch.linalg.cross(zf, xr)
        out.append(torch.stack([xr, yd, zf, pos], dim=1).float())
    return out


def make_cfg(device="cuda", sh_order=4, **overrides):
    """A reference-style cfg object (attribute access + .get) with the benchmark's settings."""
    from .utils.misc import Config

    cfg = Config(
        device=device, sh_order=sh_order, svec_act="exp", alpha_act="sigmoid",
        tile_size=16, frustum_culling_radius=1.0, tile_culling_type="aabb", tile_culling_radius=6.0,
        tile_culling_thresh=0.01, T_thresh=1e-4, warm_up=0, adaptive_control_iteration=0,
        pos_grad_thresh=2e-4, split_scale_thresh=0.01, scale_shrink_factor=1.6, alpha_reset_period=0,
        remove_low_alpha_period=0, alpha_reset_val=0.01, alpha_thresh=0.005, sh_upgrades=[],
        split_type="2d_mean_grad", split_reduction="mean", depth_detach=True, bg=False,
        skip_frustum_culling=False, max_iteration=30000, lr=1e-3, mean_lr=1.6e-4, qvec_lr=1e-3,
        svec_lr=5e-3, sh_coeffs_lr=2.5e-3, alpha_lr=5e-2, warmup_steps=0, debug=False,
        svec_init_method="fixed", svec_init=0.01, alpha_init=0.5,
    )
    cfg.update(overrides)
    return cfg


def renderer_from_scene(scene, cfg):
    """Build an SHRenderer holding the scene's parameters (device from cfg)."""
    from .gs.sh_renderer import SHRenderer

    r = SHRenderer(cfg)
    dev = cfg.device
    r._set_params({
        "mean": scene["mean"].to(dev), "qvec": scene["qvec"].to(dev),
        "svec_before_activation": scene["svec_before_activation"].to(dev),
        "sh_coeffs": scene["sh_coeffs"].to(dev),
        "alpha_before_activation": scene["alpha_before_activation"].to(dev),
    })
    r._reset_adc_buffers()
    r.now_C = scene["C"]
    return r
