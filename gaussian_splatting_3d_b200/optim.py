"""FusedAdam -- `opt.step()` of the reference's training loop (main_sh.py:193) as one kernel launch.

SHRenderer.get_optimizer (sh_renderer.py:720-729) builds `torch.optim.Adam(groups, lr, betas=(0.9, 0.99))`
with one group per parameter tensor and its scheduled learning rate; the loop calls `opt.zero_grad()`,
`loss.backward()`, `opt.step()` and then RE-CREATES the optimiser (`opt = renderer.get_optimizer(e)`,
main_sh.py:238).  This class keeps that interface (`torch.optim.Optimizer`: param_groups, state,
zero_grad, step, state_dict) and torch's arithmetic (`_single_tensor_adam`, no weight decay / amsgrad /
maximize) but runs all groups through gs3d_adam_step:

  * first step of a fresh optimiser: moments written, never read (20 B/element instead of 28);
  * `single_step=True` (opt-in: exactly what a loop that re-creates the optimiser after every step
    needs): no moment buffers at all (12 B/element, and no 2 x 708 MB allocation per step at 3 M
    Gaussians); a second step() raises.

CUDA only, FP32, contiguous; anything else raises (there is no CPU fallback).
"""
import torch

from . import ops

_MAX_SEGMENTS = 8


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, single_step=False):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: invalid lr / betas / eps")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.single_step = bool(single_step)
        self._steps_taken = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.single_step and self._steps_taken:
            raise RuntimeError("FusedAdam(single_step=True) keeps no moments and is valid for one step only; "
                               "re-create it (SHRenderer.get_optimizer) or use single_step=False")
        # bucket by (betas, eps, step, mode): one launch per bucket of <= 8 tensors
        buckets = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not g.is_contiguous():
                    g = g.contiguous()
                if self.single_step:
                    step, mode, m, v = 1, 2, None, None
                else:
                    st = self.state[p]
                    if not st:
                        st["step"] = torch.tensor(0.0)
                        st["exp_avg"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                        st["exp_avg_sq"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                        mode = 1  # the kernel writes both moments without reading them
                    else:
                        mode = 0
                    st["step"] += 1
                    step, m, v = int(st["step"].item()), st["exp_avg"], st["exp_avg_sq"]
                key = (b1, b2, group["eps"], step, mode)
                buckets.setdefault(key, []).append((p, g, m, v, group["lr"]))
        for (b1, b2, eps, step, mode), items in buckets.items():
            for s in range(0, len(items), _MAX_SEGMENTS):
                chunk = items[s:s + _MAX_SEGMENTS]
                ops.adam_step([c[0] for c in chunk], [c[1] for c in chunk], [c[2] for c in chunk],
                              [c[3] for c in chunk], [c[4] for c in chunk], b1, b2, eps, step, mode)
        self._steps_taken += 1
        return loss
