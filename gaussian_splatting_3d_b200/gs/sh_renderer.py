"""SHRenderer -- the SH model module of the reference (gs/sh_renderer.py) on the B200 kernels.

Public surface kept (sh_renderer.py:30-768): `sh_base`, `init_sh_coeffs`, `SHRenderer(cfg, pts, rgb)`
with `forward(c2w, camera_info) -> [H,W,3]`, parameters `mean, qvec, svec_before_activation,
sh_coeffs, alpha_before_activation`, properties `svec, alpha`, state `N, now_C, max_C, grad_mean,
cnt, depth, radius, total_dub_gaussians, frustum_culling_mask, mean_2d`, and the methods
`adaptive_control, update_grads, split_gaussians, remove_low_alpha_gaussians, reset_alpha, save,
load, get_param_groups, get_optimizer, select_masked_gaussians, vis_grads_gaussians, log*,
to_pointcloud`.

What changed underneath forward() (sh_renderer.py:188-316): one fused kernel replaces frustum
planes + sphere cull + five boolean-mask gathers of ALL parameters + ~45 ATen ops of projection
and rect counting; binning is the hand-written radix path; compositing and its backward are the
sm_100a kernels; the autograd graph is a single node over the leaf parameters.  Gaussian ids are
original indices (nothing is compacted), so `mean_2d`, `depth` are [N,...] with zero rows for culled
Gaussians, and `mean_2d.grad` is set to the full [N,2] gradient after backward.
"""
import gc
from pathlib import Path

import numpy as np
import torch

from ..utils.activations import activations, inv_activations
from ..utils.misc import print_info
from ..utils.schedulers import lr_schedulers
from .. import ops
from ..optim import FusedAdam
from .backend import _backend  # noqa: F401  (import fails loudly when the CUDA library is missing)
from .renderer import splat_sh, step_check

sh_base = 0.28209479177387814

_PARAM_NAMES = ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")


@torch.no_grad()
def init_sh_coeffs(cfg, rgb, C: int):
    """Degree-0 coefficients that reproduce `rgb` through sigmoid(sh * Y00) (sh_renderer.py:33-38)."""
    sh_coeffs = torch.zeros((rgb.shape[0], 3, C * C), device=rgb.device)
    sh_coeffs[:, :, 0] = inv_activations["sigmoid"](rgb) / sh_base
    return sh_coeffs


def cov_init(pts, k=3):
    """Mean squared distance to the k nearest neighbours (gs/initialize.py:6-23 via faiss
    IndexFlatL2, which returns squared L2).  faiss is not available offline; chunked torch.cdist."""
    pts = pts.float()
    out = torch.empty(pts.shape[0], device=pts.device)
    step = max(1, min(pts.shape[0], (1 << 26) // max(pts.shape[0], 1)))
    for s in range(0, pts.shape[0], step):
        d2 = torch.cdist(pts[s:s + step], pts).square()
        out[s:s + step] = d2.topk(min(k + 1, pts.shape[0]), dim=1, largest=False).values[:, 1:].mean(dim=1)
    return out


def _act_code(name, kind):
    """Activation name -> flag understood by the fused kernels (1 = exp / sigmoid, 0 = identity)."""
    if name == kind:
        return 1
    if name == "nothing":
        return 0
    return None  # anything else: handled with torch ops outside the kernel


class SHRenderer(torch.nn.Module):
    def __init__(self, cfg, pts=None, rgb=None):
        super().__init__()
        self.device = cfg.device
        self.cfg = cfg
        self.max_C = cfg.sh_order
        self.svec_act = activations[cfg.svec_act]
        self.alpha_act = activations[cfg.alpha_act]
        self.svec_inv_act = inv_activations[cfg.svec_act]
        self.alpha_inv_act = inv_activations[cfg.alpha_act]
        self._svec_code = _act_code(cfg.svec_act, "exp")
        self._alpha_code = _act_code(cfg.alpha_act, "sigmoid")
        if self._svec_code is None or self._alpha_code is None:
            raise NotImplementedError("fused kernels support svec_act in {exp, nothing} and "
                                      "alpha_act in {sigmoid, nothing}")
        if pts is not None and rgb is not None:
            self.initialize(cfg, pts, rgb)
        self.now_C = 1
        self.N_changed = False
        self.depth = None
        self.radius = None
        self.total_dub_gaussians = 0
        # Opt-in: a fixed capacity (in duplicates) for the tile lists.  The forward then never reads the duplicate
        # count back (the reference's `.item()`, gs/culling.py:33-35): no host sync, every launch sized by the
        # capacity, so a whole training step can be captured in a CUDA graph (graph.GraphedStep).  When the
        # count exceeds the capacity the lists are truncated and `overflowed()` reports it.
        self.static_capacity = None
        self._overflow = None
        self._topleft_cache = {}
        self.fuse_adc = False  # accumulate grad_mean inside the backward kernel (see update_grads)
        # Optional caller-owned leaf-gradient buffers {param name: tensor}: backward ADDS into them
        # instead of allocating (parallel.FlatGradients); use torch.autograd.grad, not .backward().
        self.grad_buffers = None
        self._state = None
        if hasattr(self, "mean"):
            self._reset_adc_buffers(register=True)
        self.set_cfg(cfg)
        self.set_scheduler(cfg)

    # ------------------------------------------------------------------ construction

    def initialize(self, cfg, pts, rgb):
        self.N = pts.shape[0]
        self.mean = torch.nn.Parameter(pts)
        qvec = torch.zeros([self.N, 4])
        qvec[..., 0] = 1.0
        self.qvec = torch.nn.Parameter(qvec)
        self.sh_coeffs = torch.nn.Parameter(init_sh_coeffs(cfg, rgb, self.max_C))
        method = cfg.get("svec_init_method", "nearest")
        if method == "fixed":
            svec = torch.full([self.N, 3], float(self.svec_inv_act(cfg.svec_init)))
        elif method == "nearest":
            init_svec = (cov_init(pts, cfg.get("nearest_k", 3)) * 10).clamp(max=0.1, min=0.01)
            svec = self.svec_inv_act(init_svec).unsqueeze(1).repeat(1, 3)
        else:
            raise NotImplementedError
        self.svec_before_activation = torch.nn.Parameter(svec)
        self.alpha_before_activation = torch.nn.Parameter(
            torch.full([self.N], float(self.alpha_inv_act(cfg.alpha_init))))

    def _reset_adc_buffers(self, register=False):
        gm = torch.zeros_like(self.mean.data[..., 0])
        cnt = torch.zeros_like(self.mean.data[..., 0], dtype=torch.int32)
        if register:
            self.register_buffer("grad_mean", gm)
            self.register_buffer("cnt", cnt)
        else:
            self.grad_mean, self.cnt = gm, cnt
        # generation of the ADC buffers: parallel.sync_adc keeps a per-generation baseline
        self._adc_epoch = getattr(self, "_adc_epoch", 0) + 1

    def set_cfg(self, cfg):
        g = cfg.get
        self.tile_size = cfg.tile_size
        self.frustum_culling_radius = cfg.frustum_culling_radius
        self.tile_culling_type = g("tile_culling_type", "aabb")
        self.tile_culling_radius = cfg.tile_culling_radius
        self.tile_culling_thresh = g("tile_culling_thresh", 0.01)
        self.T_thresh = cfg.T_thresh
        self.warm_up = g("warm_up", 1000)
        self.adaptive_control_iteration = g("adaptive_control_iteration", 0)
        self.pos_grad_thresh = g("pos_grad_thresh", 0.0002)
        self.split_scale_thresh = g("split_scale_thresh", 0.01)
        self.scale_shrink_factor = g("scale_shrink_factor", 1.6)
        self.alpha_reset_period = g("alpha_reset_period", 0)
        self.remove_low_alpha_period = g("remove_low_alpha_period", 0)
        self.alpha_reset_val = g("alpha_reset_val", 0.01)
        self.alpha_thresh = g("alpha_thresh", 0.005)
        self.remove_tiny_period = g("remove_tiny_period", 500)
        self.remove_tiny = g("remove_tiny", False)
        self.split_type = g("split_type", "mean_grad")
        self.split_reduction = g("split_reduction", "max")
        self.svec_thresh = g("svec_thresh", 50)
        self.remove_large_period = g("remove_large_period", 500)
        self.world_large_thresh = g("world_large_thresh", 30)
        self.sh_upgrades = g("sh_upgrades", [])
        self.depth_detach = g("depth_detach", True)
        self.bg = g("bg", False)
        if self.bg:
            self.bg_rgb = torch.FloatTensor(list(g("bg_rgb", [1.0, 1.0, 1.0]))).to(self.device)
        self.skip_frustum_culling = g("skip_frustum_culling", False)
        self.exact_decisions = g("exact_decisions", True)

    def set_scheduler(self, cfg):
        g = cfg.get
        self.scheduler = {}
        for name in ("mean", "svec", "qvec", "sh_coeffs", "alpha"):
            lr = g(f"{name}_lr", g("lr", 1e-3))
            self.scheduler[name] = lr_schedulers[g(f"{name}_scheduler", "nothing")](
                g("max_iteration", 1), lr, g(f"{name}_lr_end", lr),
                g(f"{name}_warmup_steps", g("warmup_steps", 0)))

    # ------------------------------------------------------------------ hot path

    def forward(self, c2w, camera_info):
        """c2w [3,4] float32 (OpenCV axes) on the model's device -> image [H,W,3]
        (sh_renderer.py:188-316)."""
        if self.tile_culling_type != "aabb":
            raise NotImplementedError("only tile_culling_type 'aabb' is on the hot path")
        training_2d = self.split_type == "2d_mean_grad" and self.training
        adc_mode = 0
        if self.fuse_adc and training_2d:
            adc_mode = 1 if self.split_reduction == "max" else 2
        state = {
            "camera_info": camera_info, "tile_size": self.tile_size, "C": self.now_C,
            "svec_act": self._svec_code, "alpha_act": self._alpha_code,
            "frustum_radius": self.frustum_culling_radius,
            "skip_frustum_culling": self.skip_frustum_culling, "tile_D": self.tile_culling_radius,
            "T_thresh": self.T_thresh, "detach_depth": self.depth_detach,
            "cnt": self.cnt if hasattr(self, "cnt") else None,
            "bg_rgb": self.bg_rgb if self.bg else None, "exact": self.exact_decisions,
            "adc_acc": self.grad_mean if adc_mode else None, "adc_mode": adc_mode,
            "grad_buffers": self.grad_buffers,
            "capacity": self.static_capacity, "topleft": self._topleft(camera_info),
        }
        out = splat_sh(self.mean, self.qvec, self.svec_before_activation, self.sh_coeffs,
                       self.alpha_before_activation, c2w, state)
        k1 = state.pop("out_k1")
        self._state = state
        self.total_dub_gaussians = state["n_dub"]  # int, or an int64 [1] device tensor in static-capacity mode
        self._overflow = state.get("overflow")
        self.depth = k1["depth"]
        self.radius = None
        if training_2d:
            self.frustum_culling_mask = k1["mask"]
            self.mean_2d = k1["mean2d"]
            state["mean2d_ref"] = self.mean_2d  # backward sets .grad on it
        return out.view(camera_info.h, camera_info.w, 3)

    def _topleft(self, cam):
        """(-cx/fx, -cy/fy) on the device, uploaded once per camera instead of once per frame."""
        key = (float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy), str(self.mean.device))
        t = self._topleft_cache.get(key)
        if t is None:
            t = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32).to(self.mean.device)
            if len(self._topleft_cache) > 64:
                self._topleft_cache.clear()
            self._topleft_cache[key] = t
        return t

    @property
    def total_dub_gaussians(self):
        """Duplicate count of the last forward (a Python int like the reference's; in static-capacity mode the
        count lives on the device and reading it here synchronises)."""
        v = self._n_dub
        return int(v.item()) if isinstance(v, torch.Tensor) else v

    @total_dub_gaussians.setter
    def total_dub_gaussians(self, v):
        self._n_dub = v

    def overflowed(self):
        """static-capacity mode: did the last forward need more duplicates than `static_capacity`?  (syncs)"""
        return bool(self._overflow is not None and int(self._overflow.item()) != 0)

    @property
    def svec(self):
        return self.svec_act(self.svec_before_activation)

    @property
    def alpha(self):
        return self.alpha_act(self.alpha_before_activation)

    # ------------------------------------------------------------------ logging (sh_renderer.py:326-421)

    @torch.no_grad()
    def log(self, writer, step):
        for fn in (self.log_depth_and_radius, self.log_bounds, self.log_info, self.log_grad_bounds,
                   self.log_n_gaussian_dub, self.log_statistics, self.log_lr):
            fn(writer, step)

    @torch.no_grad()
    def log_lr(self, writer, step):
        for name in self.get_param_groups():
            writer.add_scalar(f"lr/{name}", self.scheduler[name](step), step)

    @torch.no_grad()
    def log_depth_and_radius(self, writer, step):
        for tag, t in (("depth", self.depth), ("radius", self.radius)):
            if t is not None:
                writer.add_scalar(f"bounds/{tag}_max", t.max(), step)
                writer.add_scalar(f"bounds/{tag}_min", t.min(), step)
                writer.add_scalar(f"bounds/{tag}_mean", t.mean(), step)

    def _sh_now(self, t):
        return t[..., : self.now_C * self.now_C]

    @torch.no_grad()
    def log_bounds(self, writer, step):
        for tag, t in (("mean", self.mean), ("qvec", self.qvec), ("svec", self.svec),
                       ("color", self._sh_now(self.sh_coeffs)), ("alpha", self.alpha)):
            writer.add_scalar(f"bounds/{tag}_max", t.max(), step)
            writer.add_scalar(f"bounds/{tag}_min", t.min(), step)

    @torch.no_grad()
    def log_info(self, writer, step):
        writer.add_scalar("info/mean_mean", self.mean.abs().mean(), step)
        writer.add_scalar("info/qvec_mean", self.qvec.abs().mean(), step)
        writer.add_scalar("info/svec_mean", self.svec.abs().mean(), step)
        writer.add_scalar("info/sh_coeffs_mean", self._sh_now(self.sh_coeffs).abs().mean(), step)
        writer.add_scalar("info/alpha_mean", self.alpha.sigmoid().mean(), step)

    @torch.no_grad()
    def log_grad_bounds(self, writer, step):
        if self.mean.grad is None:
            return
        for tag, t in (("mean", self.mean.grad), ("qvec", self.qvec.grad),
                       ("svec", self.svec_before_activation.grad),
                       ("sh_coeffs", self._sh_now(self.sh_coeffs.grad)),
                       ("alpha", self.alpha_before_activation.grad)):
            writer.add_scalar(f"grad_bounds/{tag}_max", t.max(), step)
            writer.add_scalar(f"grad_bounds/{tag}_min", t.min(), step)

    def log_n_gaussian_dub(self, writer, step):
        writer.add_scalar("n_gaussian_dub", self.total_dub_gaussians, step)

    @torch.no_grad()
    def log_statistics(self, writer, epoch):
        writer.add_histogram("hists/mean", self.mean.norm(dim=-1).cpu().numpy(), epoch)
        writer.add_histogram("hists/svec", self.svec.max(dim=-1)[0].cpu().numpy(), epoch)
        writer.add_histogram("hists/alpha", self.alpha.cpu().numpy(), epoch)
        if self.mean.grad is not None:
            writer.add_histogram("hists/grad_mean", self.mean.grad.norm(dim=-1).cpu().numpy(), epoch)

    # ------------------------------------------------------------------ adaptive density control

    def _set_params(self, new):
        """Replace the five parameter tensors (sh_renderer.py:529-533 and friends)."""
        for name in _PARAM_NAMES:
            setattr(self, name, torch.nn.Parameter(new[name]))
        self.N = self.mean.shape[0]
        # caller-owned gradient buffers (parallel.FlatGradients.attach) are views sized for the OLD
        # parameters: drop them so the next backward cannot write past their end; the owner must
        # build a new FlatGradients for the new tensors
        if getattr(self, "grad_buffers", None) is not None:
            self.grad_buffers = None

    def _param_data(self):
        return {name: getattr(self, name).data for name in _PARAM_NAMES}

    def split_gaussians_by_radius(self):
        pass

    def _hot_classes(self):
        """KEEP / CLONE / SPLIT class per Gaussian (sh_renderer.py:433-456), one kernel."""
        if self.split_type not in ("mean_grad", "2d_mean_grad"):
            raise NotImplementedError
        if self.split_reduction not in ("mean", "max"):
            raise NotImplementedError
        return ops.adc_classify(self.grad_mean, self.cnt, self.split_reduction, self.pos_grad_thresh,
                                self.svec_before_activation.data, self._svec_code, self.split_scale_thresh)

    def split_gaussians(self):
        """Clone small / split large Gaussians whose accumulated positional gradient exceeds
        pos_grad_thresh (sh_renderer.py:426-540).  Output order: untouched + to-be-cloned originals,
        then the clones, then two samples per split Gaussian with scales / scale_shrink_factor.
        The reference's ~60 mask gathers / cats run as classify -> block scan -> one row mover; the
        `torch.randn(num_split * 2, 3)` draw stays a torch call so the RNG stream is unchanged."""
        assert self.mean.grad is not None, (
            "mean.grad is None while clone or split gaussians are performed according to spatial "
            "gradient of mean")
        plan, (n_stay, num_clone, num_split) = ops.adc_plan(self._hot_classes())
        print(f"Splitting Gaussians: num_split {num_split} num_clone {num_clone}")
        noise = torch.randn(num_split * 2, 3, device=self.mean.device)
        old = self._param_data()
        new = ops.adc_apply(plan, old["mean"], old["qvec"], old["svec_before_activation"], old["sh_coeffs"],
                            old["alpha_before_activation"], self._svec_code, self.scale_shrink_factor, noise)
        expected = self.N + num_split + num_clone
        self._set_params(dict(zip(_PARAM_NAMES, (new[0], new[1], new[2], new[3], new[4]))))
        assert self.N == expected
        print(f"num gaussians: {self.N}")
        del old, new
        gc.collect()

    def select_masked_gaussians(self, mask):
        """Keep the Gaussians with mask True (sh_renderer.py:731-741): scan + one fused row mover
        instead of five boolean-mask gathers."""
        mask = mask.contiguous()
        if mask.dtype != torch.bool:
            mask = mask != 0
        plan, _ = ops.adc_plan(mask)
        old = self._param_data()
        new = ops.adc_apply(plan, old["mean"], old["qvec"], old["svec_before_activation"], old["sh_coeffs"],
                            old["alpha_before_activation"], self._svec_code)
        self._set_params(dict(zip(_PARAM_NAMES, (new[0], new[1], new[2], new[3], new[4]))))

    def remove_low_alpha_gaussians(self):
        before = self.N
        cls = ops.adc_classify_alpha(self.alpha_before_activation.data, self._alpha_code, self.alpha_thresh)
        plan, _ = ops.adc_plan(cls)
        old = self._param_data()
        new = ops.adc_apply(plan, old["mean"], old["qvec"], old["svec_before_activation"], old["sh_coeffs"],
                            old["alpha_before_activation"], self._svec_code)
        self._set_params(dict(zip(_PARAM_NAMES, (new[0], new[1], new[2], new[3], new[4]))))
        print(f"remove_low_alpha_gaussians: removed {before - self.N}, remaining {self.N}")

    @torch.no_grad()
    def remove_large_gaussians(self):
        self.select_masked_gaussians((self.svec > self.world_large_thresh).all(dim=-1))

    def remove_tiny_gaussians(self):
        before = self.N
        self.select_masked_gaussians((self.svec > self.svec_tiny_thresh).all(dim=-1))
        print(f"removed {before - self.N} gaussians")

    def reset_alpha(self):
        self.alpha_before_activation.data.fill_(self.alpha_inv_act(self.alpha_reset_val))

    @torch.no_grad()
    def update_grads(self):
        """Accumulate the positional-gradient statistic used by split_gaussians
        (sh_renderer.py:602-623).  With `fuse_adc` the 2-D variant has already been accumulated
        inside the projection-backward kernel and this is a no-op for it."""
        if self.split_type == "mean_grad":
            g = self.mean.grad.norm(dim=-1)
        elif self.split_type == "2d_mean_grad":
            if self.fuse_adc:
                return
            # full-[N] gradient with zero rows for culled Gaussians == masked update of the reference
            g = self.mean_2d.grad.norm(dim=-1)
        else:
            raise NotImplementedError
        if self.split_reduction == "max":
            self.grad_mean = torch.maximum(self.grad_mean, g)
        elif self.split_reduction == "mean":
            self.grad_mean += g
        else:
            raise NotImplementedError

    def adaptive_control(self, epoch, force=False):
        """sh_renderer.py:626-661."""
        if epoch in self.sh_upgrades:
            self.now_C = min(self.now_C + 1, self.max_C)
            print(f"Spherical Harmonics now: {self.now_C}")
        if epoch < self.warm_up:
            return
        self.update_grads()
        if step_check(epoch, self.adaptive_control_iteration) or force:
            self.split_gaussians()
            self.N_changed = True
        if step_check(epoch, self.remove_low_alpha_period) or force:
            self.remove_low_alpha_gaussians()
            self.N_changed = True
        if step_check(epoch, self.alpha_reset_period) or force:
            self.reset_alpha()
            self.N_changed = True
        if self.N_changed:
            self._reset_adc_buffers()
            gc.collect()
            self.N_changed = False
            torch.cuda.empty_cache()

    # ------------------------------------------------------------------ checkpoint (sh_renderer.py:663-708)

    def save(self, path):
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        state = {name: getattr(self, name).data for name in _PARAM_NAMES}
        state["N"] = self.N
        state["cfg"] = self.cfg
        torch.save(state, path)

    @classmethod
    def load(cls, path, cfg=None):
        state = torch.load(path, weights_only=False)
        renderer = cls(cfg if cfg is not None else state["cfg"])
        renderer._set_params({name: state[name] for name in _PARAM_NAMES})
        assert renderer.N == state["N"]
        renderer._reset_adc_buffers()
        return renderer

    # ------------------------------------------------------------------ optimiser (sh_renderer.py:710-729)

    def get_param_groups(self):
        return {
            "mean": self.mean,
            "qvec": self.qvec,
            "svec": self.svec_before_activation,
            "sh_coeffs": self.sh_coeffs,
            "alpha": self.alpha_before_activation,
        }

    def get_optimizer(self, epoch):
        groups = [{"params": p, "lr": self.scheduler[name](epoch)}
                  for name, p in self.get_param_groups().items()]
        if not self.cfg.get("fused_adam", True):
            return torch.optim.Adam(groups, lr=self.cfg.lr, betas=(0.9, 0.99))
        # same interface and arithmetic, one launch for all five groups; `adam_single_step: true` is for
        # loops that, like main_sh.py:238, re-create the optimiser after every step (no moment buffers)
        return FusedAdam(groups, lr=self.cfg.lr, betas=(0.9, 0.99),
                         single_step=self.cfg.get("adam_single_step", False))

    def vis_grads_gaussians(self, thresh):
        self.select_masked_gaussians(self.mean_2d.grad.norm(dim=-1) > thresh)

    def to_pointcloud(self):
        return {
            "pos": self.mean.data.cpu().numpy(),
            "rgb": torch.sigmoid(self.sh_coeffs.data[..., 0]).cpu().numpy(),
            "alpha": self.alpha.data.cpu().numpy(),
        }
