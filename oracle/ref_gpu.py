"""The reference's SHRenderer.forward flow (gs/sh_renderer.py:188-316) driven on the GPU through the
REAL reference CUDA extension (oracle/_ref/_gs_ref*.so, built by oracle/build_ref.py from
/root/reference/gs/src).  TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py --impl reference and
by tests; none of the product's kernels are on this path -- the torch-level half is the reference's
own op sequence (oracle/ref_torch.py, pinned against the real reference by tests/golden) and the
kernels are the reference's.
"""
import importlib.util
from pathlib import Path

import torch

from . import ref_torch as R

HERE = Path(__file__).resolve().parent


def load_reference_extension():
    """-> module with the reference's 20 bindings, or None when the .so is absent / unloadable."""
    sos = sorted((HERE / "_ref").glob("_gs_ref*.so"))
    if not sos:
        return None
    spec = importlib.util.spec_from_file_location("_gs_ref", sos[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_render_fn(ext, bg_rgb=None):
    """bg_rgb: None -> _render_sh (gs/renderer.py:672-828); a float32 [3] device tensor -> _render_sh_bg
    (gs/renderer.py:831-993, the *_with_bg bindings)."""

    class _render_sh(torch.autograd.Function):
        """gs/renderer.py:672-828 (argument order, zero-initialised outputs, saved tensors)."""

        @staticmethod
        def forward(ctx, mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, consts):
            H, W = consts[5], consts[6]
            out = torch.zeros([H * W * 3], dtype=torch.float32, device=mean.device)
            if bg_rgb is None:
                ext.tile_based_vol_rendering_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                                topleft, c2w, *consts)
            else:
                ext.tile_based_vol_rendering_sh_with_bg(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                                        topleft, c2w, *consts, bg_rgb)
            ctx.save_for_backward(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w)
            ctx.consts = consts
            return out

        @staticmethod
        def backward(ctx, grad):
            mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w = ctx.saved_tensors
            grad_mean = torch.zeros_like(mean)
            grad_cov = torch.zeros_like(cov)
            grad_sh = torch.zeros_like(sh_coeffs)
            grad_alpha = torch.zeros_like(alpha)
            if bg_rgb is None:
                ext.tile_based_vol_rendering_backward_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                                         grad_mean, grad_cov, grad_sh, grad_alpha,
                                                         grad.contiguous(), topleft, c2w, *ctx.consts)
            else:
                ext.tile_based_vol_rendering_backward_sh_with_bg(
                    mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_sh,
                    grad_alpha, grad.contiguous(), topleft, c2w, *ctx.consts, bg_rgb)
            return grad_mean, grad_cov, grad_sh, grad_alpha, None, None, None, None, None, None

    return _render_sh.apply


class ReferenceGPURenderer:
    """Holds leaf parameters on the GPU and renders like the reference's SHRenderer."""

    def __init__(self, ext, scene, device, C, tile_size=16, frustum_radius=1.0, tile_D=6.0, T_thresh=1e-4,
                 bg_rgb=None):
        self.ext = ext
        self.render = make_render_fn(ext, bg_rgb)
        self.aux = None  # intermediate tensors of the last forward (for stage-by-stage comparison)
        self.dev = device
        self.C = C
        self.tile_size, self.frustum_radius, self.tile_D, self.T_thresh = tile_size, frustum_radius, tile_D, T_thresh
        names = ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")
        self.params = {k: scene[k].to(device).clone().requires_grad_(True) for k in names}
        self.total_dub_gaussians = 0

    def zero_grad(self):
        for p in self.params.values():
            p.grad = None

    def forward(self, c2w, cam):
        p, ext, dev = self.params, self.ext, self.dev
        C, tile = self.C, self.tile_size
        f_normals, f_pts = R.get_frustum(c2w, cam)
        N = p["mean"].shape[0]
        mask = torch.zeros(N, dtype=torch.bool, device=dev)
        svec = torch.exp(p["svec_before_activation"])
        alpha_act = torch.sigmoid(p["alpha_before_activation"])
        with torch.no_grad():
            ext.culling_gaussian_bsphere(p["mean"], p["qvec"], svec, f_normals.contiguous(), f_pts.contiguous(),
                                         mask, self.frustum_radius)
        mean = p["mean"][mask].contiguous()
        qvec = p["qvec"][mask].contiguous()
        svec_m = svec[mask].contiguous()
        sh = p["sh_coeffs"][mask].contiguous()
        alpha = alpha_act[mask].contiguous()
        mean2d, cov, JW, depth = R.project_gaussians(mean, qvec, svec_m, c2w, True)
        if mean2d.requires_grad:
            mean2d.retain_grad()  # sh_renderer.py:217-221
        n_dub, tl, br = R.tile_culling_aabb_count(mean2d, cov, tile, cam, self.tile_D)
        self.total_dub_gaussians = n_dub
        H, W = cam.h, cam.w
        nth = H // tile + (H % tile > 0)
        ntw = W // tile + (W % tile > 0)
        topleft = torch.FloatTensor([-cam.cx / cam.fx, -cam.cy / cam.fy]).to(dev)
        start = -torch.ones([nth * ntw], dtype=torch.int32, device=dev)
        end = -torch.ones([nth * ntw], dtype=torch.int32, device=dev)
        ids = torch.zeros([n_dub], dtype=torch.int32, device=dev)
        ext.tile_culling_aabb_start_end(tl, br, ids, start, end, depth, nth, ntw)
        consts = (tile, nth, ntw, 1.0 / cam.fx, 1.0 / cam.fy, H, W, C, self.T_thresh)
        out = self.render(mean2d, cov, sh[..., : C * C].contiguous(), alpha, start, end, ids, topleft, c2w, consts)
        self.aux = dict(mask=mask, mean2d=mean2d, cov=cov, depth=depth, tl=tl, br=br, n_dub=n_dub, ids=ids,
                        start=start, end=end)
        return out.view(H, W, 3)
