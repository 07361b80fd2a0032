"""Torch (CPU, FP32) restatement of the reference's PyTorch-level half of the hot path, plus the
whole SHRenderer.forward flow glued to the C oracle kernels.  TEST INFRASTRUCTURE ONLY.

Each function follows the reference op for op (same torch calls in the same order), so on the
same machine it reproduces the reference bit for bit; tests/golden/ holds outputs of the REAL
reference functions (imported from /root/reference by oracle/make_golden.py) that pin this file.

  get_frustum               utils/camera.py:249-283
  quaternion_to_rotation_matrix   kornia 0.6.x (requirements.txt:5 `kornia`, unpinned; the
                            `QuaternionCoeffOrder.WXYZ` call sites utils/transforms.py:34-36 only exist
                            in 0.6.x): normalize_quaternion (F.normalize p=2 eps=1e-12) then the
                            tx/ty/tz product form.
  qsvec2rotmat_batched      utils/transforms.py:31-45
  project_pts / jacobian / project_gaussians   gs/renderer.py:366-419
  camera_space_to_pixel_space   utils/camera.py:290-303
  tile_culling_aabb_count   gs/culling.py:8-37
  reference_forward         gs/sh_renderer.py:188-316 (+ activations :318-324)
  split_gaussians / select_masked_gaussians / remove_low_alpha_mask   gs/sh_renderer.py:426-560,731-741
  adam_first_step           torch.optim.Adam as main_sh.py:193,238 uses it (re-created every step)
  ssim / ssim_loss / get_loss_fn   utils/loss.py:5-24 over kornia 0.6.x `kornia.losses.ssim_loss` (kornia is
                            un-vendored and unpinned, requirements.txt:5: PARITY UNPINNED at this boundary --
                            no reference test or fixture holds an SSIM value; the published algorithm is
                            restated and pinned by closed-form cases only, tests/test_loss.py)
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import gs_oracle as K


def get_frustum(c2w, cam):
    up = -c2w[:, 1]
    right = c2w[:, 0]
    lookat = c2w[:, 2]
    t = c2w[:, 3]
    yfov = 2 * np.arctan(cam.h / (2 * cam.fy))
    aspect = cam.w / cam.h
    half_vside = cam.far_plane * np.tan(yfov * 0.5)
    half_hside = half_vside * aspect
    near_point = cam.near_plane * lookat
    far_point = cam.far_plane * lookat
    near_normal = lookat
    far_normal = -lookat
    left_normal = torch.linalg.cross(far_point - half_hside * right, up)
    right_normal = torch.linalg.cross(up, far_point + half_hside * right)
    up_normal = torch.linalg.cross(far_point + half_vside * up, right)
    down_normal = torch.linalg.cross(right, far_point - half_vside * up)
    pts = torch.stack([near_point + t, far_point + t, t, t, t, t], dim=0)
    normals = torch.stack([near_normal, far_normal, left_normal, right_normal, up_normal, down_normal], dim=0)
    normals = F.normalize(normals, dim=-1)
    return normals, pts


def quaternion_to_rotation_matrix(quaternion):
    q = F.normalize(quaternion, p=2.0, dim=-1, eps=1e-12)
    w, x, y, z = torch.chunk(q, chunks=4, dim=-1)
    tx = 2.0 * x
    ty = 2.0 * y
    tz = 2.0 * z
    twx = tx * w
    twy = ty * w
    twz = tz * w
    txx = tx * x
    txy = ty * x
    txz = tz * x
    tyy = ty * y
    tyz = tz * y
    tzz = tz * z
    one = torch.tensor(1.0)
    matrix = torch.stack(
        (one - (tyy + tzz), txy - twz, txz + twy,
         txy + twz, one - (txx + tzz), tyz - twx,
         txz - twy, tyz + twx, one - (txx + tyy)), dim=-1).view(-1, 3, 3)
    return matrix


def qsvec2rotmat_batched(qvec, svec):
    return svec.unsqueeze(-2) * quaternion_to_rotation_matrix(qvec)


@torch.no_grad()
def jacobian(u):
    l = torch.norm(u, dim=-1)
    J = torch.zeros(u.size(0), 3, 3).to(u)
    J[..., 0, 0] = 1.0 / u[..., 2]
    J[..., 2, 0] = u[..., 0] / l
    J[..., 1, 1] = 1.0 / u[..., 2]
    J[..., 2, 1] = u[..., 1] / l
    J[..., 0, 2] = -u[..., 0] / u[..., 2] / u[..., 2]
    J[..., 1, 2] = -u[..., 1] / u[..., 2] / u[..., 2]
    J[..., 2, 2] = u[..., 2] / l
    return J


def project_pts(pts, c2w):
    d = -c2w[..., :3, 3]
    W = torch.transpose(c2w[..., :3, :3], -1, -2)
    return torch.einsum("ij,bj->bi", W, pts + d)


def project_gaussians(mean, qvec, svec, c2w, detach_depth=True):
    projected_mean = project_pts(mean, c2w)
    rotmat = qsvec2rotmat_batched(qvec, svec)
    sigma = rotmat @ torch.transpose(rotmat, -1, -2)
    W = torch.transpose(c2w[:3, :3], -1, -2)
    J = jacobian(projected_mean)
    JW = torch.einsum("bij,jk->bik", J, W)
    projected_cov = torch.bmm(torch.bmm(JW, sigma), torch.transpose(JW, -1, -2))[..., :2, :2].contiguous()
    if detach_depth:
        depth = projected_mean[..., 2:].clone().contiguous().detach()
    else:
        depth = projected_mean[..., 2:].clone().contiguous()
    projected_mean = projected_mean[..., :2].contiguous() / depth
    return projected_mean, projected_cov, JW, depth


def camera_space_to_pixel_space(cam, pts):
    pts[:, 0] = pts[:, 0] * cam.fx + cam.cx
    pts[:, 1] = pts[:, 1] * cam.fy + cam.cy
    return pts.to(torch.int32)


@torch.no_grad()
def tile_culling_aabb_count(mean, cov, tile_size, cam, D):
    aabb_x = torch.sqrt(D * cov[:, 0, 0])
    aabb_y = torch.sqrt(D * cov[:, 1, 1])
    side = torch.stack([aabb_x, aabb_y], dim=-1)
    tl = camera_space_to_pixel_space(cam, mean - side)
    br = camera_space_to_pixel_space(cam, mean + side)
    tl[..., 0].clamp_(min=0, max=cam.w - 1)
    tl[..., 1].clamp_(min=0, max=cam.h - 1)
    br[..., 0].clamp_(min=0, max=cam.w - 1)
    br[..., 1].clamp_(min=0, max=cam.h - 1)
    tl = torch.div(tl, tile_size, rounding_mode="floor")
    br = torch.div(br, tile_size, rounding_mode="floor")
    n = torch.prod(br - tl + 1, dim=-1).sum().item()
    return n, tl, br


class _RenderSH(torch.autograd.Function):
    """gs/renderer.py:672-828 with the C oracle standing in for the CUDA bindings."""

    @staticmethod
    def forward(ctx, mean, cov, sh, alpha, start, end, ids, topleft, c2w, consts, bg):
        out = K.render_sh_forward(mean.detach().numpy(), cov.detach().numpy(), sh.detach().numpy(),
                                  alpha.detach().numpy(), start, end, ids, topleft, c2w.numpy(), *consts,
                                  bg_rgb=bg)
        ctx.save_for_backward(mean, cov, sh, alpha)
        ctx.misc = (start, end, ids, topleft, c2w, consts, out)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad):
        mean, cov, sh, alpha = ctx.saved_tensors
        start, end, ids, topleft, c2w, consts, out = ctx.misc
        gm, gc, gs, ga = K.render_sh_backward(mean.detach().numpy(), cov.detach().numpy(),
                                              sh.detach().numpy(), alpha.detach().numpy(), start, end, ids,
                                              out, grad.contiguous().numpy(), topleft, c2w.numpy(), *consts)
        return (torch.from_numpy(gm), torch.from_numpy(gc).view_as(cov), torch.from_numpy(gs),
                torch.from_numpy(ga), None, None, None, None, None, None, None)


def reference_forward(params, c2w, cam, C, tile_size=16, frustum_radius=1.0, tile_D=6.0, T_thresh=1e-4,
                      depth_detach=True, skip_frustum_culling=False, bg_rgb=None, return_aux=False):
    """The reference's SHRenderer.forward on CPU.  `params`: dict of leaf tensors (requires_grad ok):
    mean, qvec, svec_before_activation, sh_coeffs [N,3,maxC^2], alpha_before_activation."""
    svec = torch.exp(params["svec_before_activation"])
    alpha_act = torch.sigmoid(params["alpha_before_activation"])
    N = params["mean"].shape[0]
    if skip_frustum_culling:
        mask = torch.ones(N, dtype=torch.bool)
    else:
        f_normals, f_pts = get_frustum(c2w, cam)
        mask = torch.from_numpy(K.culling_gaussian_bsphere(
            params["mean"].detach().numpy(), svec.detach().numpy(), f_normals.numpy(), f_pts.numpy(),
            frustum_radius))
    mean = params["mean"][mask].contiguous()
    qvec = params["qvec"][mask].contiguous()
    svec_m = svec[mask].contiguous()
    sh = params["sh_coeffs"][mask].contiguous()
    alpha = alpha_act[mask].contiguous()
    mean2d, cov, JW, depth = project_gaussians(mean, qvec, svec_m, c2w, depth_detach)
    if mean2d.requires_grad:
        mean2d.retain_grad()
    n_dub, tl, br = tile_culling_aabb_count(mean2d, cov, tile_size, cam, tile_D)
    H, W = cam.h, cam.w
    nth = H // tile_size + (H % tile_size > 0)
    ntw = W // tile_size + (W % tile_size > 0)
    topleft = torch.FloatTensor([-cam.cx / cam.fx, -cam.cy / cam.fy]).numpy()
    ids, start, end, keys = K.tile_culling_aabb_start_end(tl.numpy(), br.numpy(), depth.detach().numpy(),
                                                          n_dub, nth, ntw)
    psx, psy = 1.0 / cam.fx, 1.0 / cam.fy
    consts = (tile_size, nth, ntw, np.float32(psx), np.float32(psy), H, W, C, T_thresh)
    out = _RenderSH.apply(mean2d, cov, sh[..., : C * C].contiguous(), alpha, start, end, ids, topleft,
                          c2w, consts, None if bg_rgb is None else np.asarray(bg_rgb, dtype=np.float32))
    img = out.view(H, W, 3)
    if return_aux:
        return img, dict(mask=mask, mean2d=mean2d, cov=cov, depth=depth, tl=tl, br=br, n_dub=n_dub,
                         ids=ids, start=start, end=end, keys=keys, alpha=alpha, sh=sh, JW=JW)
    return img


# ---------------------------------------------------------------- adaptive density control (8f rank 1)

def split_masks(grad_mean, cnt, svec, split_reduction, pos_grad_thresh, split_scale_thresh):
    """sh_renderer.py:433-456 -> (split_mask, clone_mask)."""
    if split_reduction == "mean":
        mask = grad_mean / (cnt + 1e-5) > pos_grad_thresh
    elif split_reduction == "max":
        mask = grad_mean > pos_grad_thresh
    else:
        raise NotImplementedError
    svec_mask = (svec > split_scale_thresh).any(dim=-1)
    split_mask = torch.logical_and(mask, svec_mask)
    clone_mask = torch.logical_and(mask, torch.logical_not(split_mask))
    return split_mask, clone_mask


def split_gaussians(params, grad_mean, cnt, split_reduction, pos_grad_thresh, split_scale_thresh,
                    scale_shrink_factor, noise=None, svec_act=torch.exp, svec_inv_act=torch.log):
    """sh_renderer.py:426-540 op for op.  params: dict of the five tensors (mean, qvec,
    svec_before_activation, sh_coeffs, alpha_before_activation).  `noise` replaces the
    `torch.randn(num_split * 2, 3)` draw of :470 (None: draw it here, like the reference).
    -> (new params dict, num_split, num_clone)."""
    mean, qvec = params["mean"], params["qvec"]
    svec_ba, sh, alpha_ba = params["svec_before_activation"], params["sh_coeffs"], params["alpha_before_activation"]
    svec = svec_act(svec_ba)
    split_mask, clone_mask = split_masks(grad_mean, cnt, svec, split_reduction, pos_grad_thresh, split_scale_thresh)
    num_split = int(split_mask.sum().item())
    num_clone = int(clone_mask.sum().item())
    split_mean = mean[split_mask].repeat(2, 1)
    split_qvec = qvec[split_mask].repeat(2, 1)
    split_svec = svec[split_mask].repeat(2, 1)
    split_sh = sh[split_mask].repeat(2, 1, 1)
    split_alpha = alpha_ba[split_mask].repeat(2)
    split_rotmat = quaternion_to_rotation_matrix(split_qvec).transpose(-1, -2)
    if noise is None:
        noise = torch.randn(num_split * 2, 3, device=mean.device)
    split_gn = noise * split_svec
    split_sampled_mean = split_mean + torch.einsum("bij, bj -> bi", split_rotmat, split_gn)
    N_old = mean.shape[0]
    unchanged = N_old - num_split
    N = N_old + num_split + num_clone
    new = {
        "mean": torch.zeros([N, 3]), "qvec": torch.zeros([N, 4]), "svec_before_activation": torch.zeros([N, 3]),
        "sh_coeffs": torch.zeros([N] + list(sh.shape[1:])), "alpha_before_activation": torch.zeros([N]),
    }
    keep = ~split_mask
    for k in new:
        new[k][:unchanged] = params[k][keep]
        new[k][unchanged:unchanged + num_clone] = params[k][clone_mask]
    pts = unchanged + num_clone
    new["mean"][pts:pts + 2 * num_split] = split_sampled_mean
    new["qvec"][pts:pts + 2 * num_split] = split_qvec
    new["sh_coeffs"][pts:pts + 2 * num_split] = split_sh
    new["alpha_before_activation"][pts:pts + 2 * num_split] = split_alpha
    new["svec_before_activation"][pts:pts + 2 * num_split] = svec_inv_act(split_svec / scale_shrink_factor)
    assert pts + 2 * num_split == N
    return new, num_split, num_clone


def select_masked_gaussians(params, mask):
    """sh_renderer.py:731-741 (also :542-560, :590-600): boolean-mask gather of the five tensors."""
    return {k: v[mask] for k, v in params.items()}


def remove_low_alpha_mask(alpha_ba, alpha_thresh, alpha_act=torch.sigmoid):
    """sh_renderer.py:543."""
    return alpha_act(alpha_ba) >= alpha_thresh


def adam_first_step(params, grads, lrs, betas=(0.9, 0.99), eps=1e-8):
    """One step of a freshly created torch.optim.Adam (the only kind main_sh.py ever takes: the optimiser
    is re-created after every step, main_sh.py:238), with sh_renderer.py:720-729's per-tensor lrs.
    Uses torch.optim.Adam itself; params are updated in place."""
    ps = [torch.nn.Parameter(p) for p in params]
    for p, g in zip(ps, grads):
        p.grad = g
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)], lr=1e-3, betas=betas, eps=eps)
    opt.step()
    return [p.data for p in ps], opt


# ---------------------------------------------------------------- loss (8f rank 4)

def gaussian_window(window_size, sigma=1.5):
    """kornia.filters.kernels.gaussian: exp(-x^2 / (2 sigma^2)) normalised (FP32)."""
    x = torch.arange(window_size, dtype=torch.float32) - window_size // 2
    if window_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def _filter2d_separable(img, k1d):
    """kornia.filters.filter2d_separable(img, k, k, border_type="reflect") for img [B,C,H,W]."""
    r = k1d.numel() // 2
    c = img.shape[1]
    x = F.pad(img, (r, r, r, r), mode="reflect")
    x = F.conv2d(x, k1d.view(1, 1, 1, -1).expand(c, 1, 1, -1), groups=c)
    return F.conv2d(x, k1d.view(1, 1, -1, 1).expand(c, 1, -1, 1), groups=c)


def ssim(img1, img2, window_size, max_val=1.0, eps=1e-12):
    """kornia.metrics.ssim (0.6.x, padding="same"): img [B,C,H,W] -> ssim map [B,C,H,W]."""
    k = gaussian_window(window_size).to(img1)
    C1, C2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mu1, mu2 = _filter2d_separable(img1, k), _filter2d_separable(img2, k)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    sigma1_sq = _filter2d_separable(img1 ** 2, k) - mu1_sq
    sigma2_sq = _filter2d_separable(img2 ** 2, k) - mu2_sq
    sigma12 = _filter2d_separable(img1 * img2, k) - mu1_mu2
    num = (2.0 * mu1_mu2 + C1) * (2.0 * sigma12 + C2)
    den = (mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2)
    return num / (den + eps)


def ssim_loss(img1, img2, window_size, max_val=1.0, eps=1e-12, reduction="mean"):
    """kornia.losses.ssim_loss: clamp((1 - ssim) / 2, 0, 1), mean."""
    loss = torch.clamp((1.0 - ssim(img1, img2, window_size, max_val, eps)) / 2, min=0, max=1)
    return loss.mean() if reduction == "mean" else (loss.sum() if reduction == "sum" else loss)


def get_loss_fn(loss_name, ssim_loss_mult, ssim_loss_win_size):
    """utils/loss.py:5-24 op for op (out, gt are [H,W,3])."""
    base = {"l2": F.mse_loss, "l1": F.l1_loss}[loss_name]

    def loss_fn(out, gt):
        return ssim_loss_mult * ssim_loss(out.moveaxis(-1, 0).unsqueeze(0), gt.moveaxis(-1, 0).unsqueeze(0),
                                          ssim_loss_win_size, reduction="mean") + (1 - ssim_loss_mult) * base(out, gt)

    return loss_fn
