"""Loss functions of the training loop (reference utils/loss.py:5-24) on one fused kernel sequence.

`get_loss_fn(cfg)` keeps the reference's interface: it returns `loss_fn(out, gt)` with
    loss = cfg.ssim_loss_mult * ssim_loss(out, gt, cfg.ssim_loss_win_size) + (1 - mult) * base(out, gt)
(base = mse_loss for `loss_fn: l2`, l1_loss for `l1`).  The value and d loss / d out come from
gs3d_image_loss (three HBM-bound passes) instead of kornia's ~60 ATen launches; kornia itself is not
needed (it is un-vendored and unpinned in the reference, requirements.txt:5).  CUDA only.
"""
import torch

from .. import ops


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, gt, base, mult, win):
        loss, grad = ops.image_loss(out.contiguous(), gt.contiguous(), base, mult, win,
                                    want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


def image_loss(out, gt, base="l2", ssim_loss_mult=0.0, ssim_loss_win_size=11):
    """out, gt: [H, W, 3] float32 CUDA tensors -> scalar loss (differentiable w.r.t. out)."""
    return _ImageLoss.apply(out, gt.to(out.dtype), base, float(ssim_loss_mult), int(ssim_loss_win_size))


def get_loss_fn(cfg):
    if cfg.loss_fn not in ("l1", "l2"):
        raise NotImplementedError
    base, mult, win = cfg.loss_fn, cfg.get("ssim_loss_mult", 0.0), cfg.get("ssim_loss_win_size", 11)

    def loss_fn(out, gt):
        return image_loss(out, gt, base, mult, win)

    return loss_fn
