"""Torch-level renderer ops of the reference (gs/renderer.py) on the B200 kernels.

Kept API: step_check, jacobian, project_pts, project_gaussians, render_sh, render_sh_bg
(positional argument orders of renderer.py:674-694, 833-854), plus `splat_sh`, the fused
whole-path autograd Function used by SHRenderer.forward.

Differences in mechanism, not in results:
  * project_gaussians is ONE kernel forward and ONE kernel backward (a custom autograd Function)
    instead of ~25 ATen launches + 2 bmm + 2 einsum and their autograd graph.
  * the render Functions keep the reference's contract (caller-visible tensors, in-place kernels)
    but run on the current stream with no cudaProfilerStart/Stop brackets.
The legacy RGB Functions `render` / `render_start_end` and the `Renderer` / `GaussianRenderer`
classes (renderer.py:33-362, 422-670, 1002-1549) are outside the SH hot path
and raise NotImplementedError.
"""
import torch

from .. import ops
from ..utils.misc import step_check  # noqa: F401  (re-exported like the reference)
from .backend import _backend
from .culling import tile_culling_aabb_count  # noqa: F401


@torch.no_grad()
def jacobian(u):
    """EWA Jacobian rows for camera-space points u [N,3] (renderer.py:366-377); a constant w.r.t.
    autograd (quirk Q6)."""
    l = torch.norm(u, dim=-1)
    J = torch.zeros(u.size(0), 3, 3).to(u)
    inv_z = 1.0 / u[..., 2]
    J[..., 0, 0] = inv_z
    J[..., 1, 1] = inv_z
    J[..., 0, 2] = -u[..., 0] / u[..., 2] / u[..., 2]
    J[..., 1, 2] = -u[..., 1] / u[..., 2] / u[..., 2]
    J[..., 2, 0] = u[..., 0] / l
    J[..., 2, 1] = u[..., 1] / l
    J[..., 2, 2] = u[..., 2] / l
    return J


def project_pts(pts, c2w):
    """World -> camera space, W (p - t) with W = R^T (renderer.py:381-387)."""
    d = -c2w[..., :3, 3]
    W = torch.transpose(c2w[..., :3, :3], -1, -2)
    return torch.einsum("ij,bj->bi", W, pts + d)


class _project_gaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, qvec, svec, c2w, detach_depth):
        mean_c, qvec_c, svec_c = mean.contiguous(), qvec.contiguous(), svec.contiguous()
        c2w_c = c2w.contiguous()
        mean2d, cov, JW, depth = ops.project_gaussians_forward(mean_c, qvec_c, svec_c, c2w_c, True)
        ctx.save_for_backward(mean_c, qvec_c, svec_c, c2w_c)
        ctx.detach_depth = bool(detach_depth)
        ctx.mark_non_differentiable(JW)
        if detach_depth:
            ctx.mark_non_differentiable(depth)
        return mean2d, cov, JW, depth

    @staticmethod
    def backward(ctx, g_mean2d, g_cov, g_JW, g_depth):
        mean, qvec, svec, c2w = ctx.saved_tensors
        N = mean.size(0)
        if g_mean2d is None:
            g_mean2d = torch.zeros(N, 2, dtype=torch.float32, device=mean.device)
        if g_cov is None:
            g_cov = torch.zeros(N, 2, 2, dtype=torch.float32, device=mean.device)
        gd = None if (ctx.detach_depth or g_depth is None) else g_depth
        gm, gq, gs = ops.project_gaussians_backward(mean, qvec, svec, c2w, g_mean2d, g_cov, gd,
                                                    ctx.detach_depth)
        return gm, gq, gs, None, None


def project_gaussians(mean, qvec, svec, c2w, detach_depth: bool = True):
    """-> (mean2d [N,2], cov [N,2,2], JW [N,3,3], depth [N,1]); renderer.py:391-419.
    mean2d = xy / depth with depth detached when detach_depth (the reference's "HUGE CAUTION")."""
    if not mean.is_cuda:
        raise RuntimeError("project_gaussians: CUDA tensors required (no CPU fallback)")
    return _project_gaussians.apply(mean, qvec, svec, c2w, detach_depth)


class _render_sh(torch.autograd.Function):
    """renderer.py:672-828: SH compositing through the `_gs` bindings."""

    @staticmethod
    def forward(ctx, mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, tile_size,
                n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb=None):
        out = torch.zeros([H * W * 3], dtype=torch.float32, device=mean.device)
        mean, cov, sh_coeffs, alpha = (mean.contiguous(), cov.contiguous(), sh_coeffs.contiguous(),
                                       alpha.contiguous())
        c2w = c2w.contiguous()
        consts = (tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh)
        if bg_rgb is None:
            _backend.tile_based_vol_rendering_sh(mean, cov, sh_coeffs, alpha, start, end,
                                                 gaussian_ids, out, topleft, c2w, *consts)
        else:
            _backend.tile_based_vol_rendering_sh_with_bg(mean, cov, sh_coeffs, alpha, start, end,
                                                         gaussian_ids, out, topleft, c2w, *consts,
                                                         bg_rgb)
        ctx.save_for_backward(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w)
        ctx.const = consts
        ctx.bg = bg_rgb
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w = ctx.saved_tensors
        grad_mean = torch.zeros_like(mean)
        grad_cov = torch.zeros_like(cov)
        grad_sh_coeffs = torch.zeros_like(sh_coeffs)
        grad_alpha = torch.zeros_like(alpha)
        grad = grad.contiguous()
        if ctx.bg is None:
            _backend.tile_based_vol_rendering_backward_sh(
                mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov,
                grad_sh_coeffs, grad_alpha, grad, topleft, c2w, *ctx.const)
        else:
            _backend.tile_based_vol_rendering_backward_sh_with_bg(
                mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov,
                grad_sh_coeffs, grad_alpha, grad, topleft, c2w, *ctx.const, ctx.bg)
        return (grad_mean, grad_cov, grad_sh_coeffs, grad_alpha) + (None,) * 15


def render_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, tile_size,
              n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh):
    return _render_sh.apply(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w,
                            tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C,
                            thresh, None)


def render_sh_bg(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, tile_size,
                 n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb):
    return _render_sh.apply(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w,
                            tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C,
                            thresh, bg_rgb)


# ---------------------------------------------------------------- fused whole-path Function


class _splat_sh(torch.autograd.Function):
    """The whole hot path as one autograd node over the LEAF parameters: fused cull + projection +
    rects (K1), binning (K2), compositing (K3); backward = compositing backward (K4a) + fused
    projection/activation backward (K4b).  No mask compaction, one 8-byte host read-back
    (the duplicate count), everything on the current stream."""

    @staticmethod
    def forward(ctx, mean, qvec, svec_param, sh_coeffs, alpha_param, c2w, state):
        # `state` is a plain dict of Python scalars / tensors / objects prepared by SHRenderer
        cam = state["camera_info"]
        tile = state["tile_size"]
        C = state["C"]
        dev = mean.device
        mean_c, qvec_c = mean.contiguous(), qvec.contiguous()
        svec_c, alpha_c = svec_param.contiguous(), alpha_param.contiguous()
        c2w_c = c2w.contiguous().float()
        k1 = ops.project_cull_fused(
            mean_c, qvec_c, svec_c, alpha_c, state["svec_act"], state["alpha_act"], c2w_c, cam,
            state["frustum_radius"], state["skip_frustum_culling"], state["tile_D"], tile,
            cnt=state.get("cnt"), want_records=True, want_activated=False)
        H, W = cam.h, cam.w
        nth = H // tile + (H % tile > 0)
        ntw = W // tile + (W % tile > 0)
        n_tiles = nth * ntw
        n_dub = k1["n_dub"]
        ids = torch.empty(n_dub, dtype=torch.int32, device=dev)
        start = torch.empty(n_tiles, dtype=torch.int32, device=dev)
        end = torch.empty(n_tiles, dtype=torch.int32, device=dev)
        ops.tile_culling_aabb_start_end(k1["tl"], k1["br"], ids, start, end, k1["depth"], nth, ntw,
                                        check_count=False)
        topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32).to(dev)
        psx, psy = 1.0 / cam.fx, 1.0 / cam.fy
        bg = state.get("bg_rgb")
        out = torch.zeros(H * W * 3, dtype=torch.float32, device=dev) if bg is None else \
            torch.empty(H * W * 3, dtype=torch.float32, device=dev)
        sh_c = sh_coeffs if sh_coeffs.stride(2) == 1 else sh_coeffs.contiguous()
        ops.composite_sh_forward(k1["records"], sh_c, start, end, ids, out, topleft, c2w_c, tile,
                                 nth, ntw, psx, psy, H, W, C, state["T_thresh"], bg_rgb=bg,
                                 exact=state.get("exact", True))
        ctx.save_for_backward(mean_c, qvec_c, svec_c, sh_c, alpha_c, c2w_c, k1["records"], k1["mask"],
                              ids, start, end, out, topleft)
        ctx.meta = (tile, nth, ntw, psx, psy, H, W, C, state["T_thresh"], state["svec_act"],
                    state["alpha_act"], state["detach_depth"], state.get("exact", True))
        ctx.state = state
        state["out_k1"] = k1
        state["n_dub"] = n_dub
        state["start"], state["end"], state["gaussian_ids"] = start, end, ids
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (mean, qvec, svec_p, sh, alpha_p, c2w, records, mask, ids, start, end, out,
         topleft) = ctx.saved_tensors
        tile, nth, ntw, psx, psy, H, W, C, thresh, svec_act, alpha_act, detach, exact = ctx.meta
        N = mean.size(0)
        dev = mean.device
        g_mean2d = torch.zeros(N, 2, dtype=torch.float32, device=dev)
        g_cov = torch.zeros(N, 4, dtype=torch.float32, device=dev)
        g_alpha = torch.zeros(N, dtype=torch.float32, device=dev)
        st = ctx.state
        bufs = st.get("grad_buffers")  # caller-owned leaf-gradient buffers (views of a flat buffer)
        if bufs is not None:
            g_sh = bufs["sh_coeffs"]  # accumulated into: zeroed by the owner once per step
            leaf_out = (bufs["mean"], bufs["qvec"], bufs["svec_before_activation"],
                        bufs["alpha_before_activation"])
        else:
            g_sh = torch.zeros(sh.shape, dtype=torch.float32, device=dev)
            leaf_out = None
        ops.composite_sh_backward(records, sh, start, end, ids, out, grad_out.contiguous().view(-1),
                                  g_mean2d, g_cov, g_sh, g_alpha, topleft, c2w, tile, nth, ntw, psx,
                                  psy, H, W, C, thresh, exact=exact,
                                  peer_ptrs=bufs.get("sh_peer_ptrs") if bufs else None,
                                  multicast_ptr=bufs.get("sh_multicast_ptr") if bufs else None,
                                  touched=bufs.get("touched") if bufs else None)
        if bufs is not None and bufs.get("after_composite_backward") is not None:
            bufs["after_composite_backward"]()  # e.g. the data-parallel mark broadcast, on a side stream
        gm, gq, gs, ga = ops.project_backward_fused(
            mask, mean, qvec, svec_p, alpha_p, svec_act, alpha_act, c2w, detach, g_mean2d, g_cov,
            g_alpha, grad_mean_acc=st.get("adc_acc"), adc_mode=st.get("adc_mode", 0), out=leaf_out,
            accumulate=leaf_out is not None)
        st["grad_mean2d"] = g_mean2d
        ref = st.get("mean2d_ref")
        if ref is not None:  # sh_renderer.py:217-221 `mean_2d.retain_grad()` equivalent
            ref.grad = g_mean2d
        return gm, gq, gs, g_sh, ga, None, None


def splat_sh(mean, qvec, svec_param, sh_coeffs, alpha_param, c2w, state):
    return _splat_sh.apply(mean, qvec, svec_param, sh_coeffs, alpha_param, c2w, state)


def _legacy(name):
    def fn(*a, **k):
        raise NotImplementedError(f"gs.renderer.{name}: legacy RGB path, outside the SH hot path "
                                  "(SURVEY.md 8f rank 2); use render_sh / SHRenderer.")
    fn.__name__ = name
    return fn


render = _legacy("render")
render_start_end = _legacy("render_start_end")


class GaussianRenderer(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("GaussianRenderer (legacy RGB module, renderer.py:1002-1549) is "
                                  "outside the SH hot path; use gs.sh_renderer.SHRenderer.")


Renderer = GaussianRenderer
