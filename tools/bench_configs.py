#!/usr/bin/env python
"""Side benchmarks for the other BASELINE configs (evidence for SURVEY.md 8d rows; bench.py stays the
contract bench for cfg 2 / cfg 4):

  cfg3  500 k Gaussians, C=3, 1008x756: FULL training step = forward + L2 loss + backward +
        ADC position-grad accumulation (fused into the projection-backward kernel) + cnt + Adam.
  cfg4  cfg-2 scene, 8 cameras on a ring per step, views sharded over the ranks, gradients exchanged
        (argv[3]: pull | push | fused | sparse | allreduce); strong scaling: T_G for the same 8-camera step.
  cfg5  6 M Gaussians, C=4, 3840x2160 forward; on N ranks (torchrun) tile rows are sharded.

  python tools/bench_configs.py cfg3
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py cfg5
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gaussian_splatting_3d_b200 import parallel as P  # noqa: E402
from gaussian_splatting_3d_b200 import synthetic as S  # noqa: E402


def timed(fn, k, world, dev):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_loss():
    """Loss step at cfg-2 resolution: fused kernels vs the same loss composed from torch ops on the GPU
    (what kornia's ssim_loss + F.mse_loss launch), forward + backward to d loss / d out."""
    import torch.nn.functional as F

    from gaussian_splatting_3d_b200.utils.loss import image_loss

    def torch_loss_fn(out, gt, mult=0.2, win=11):
        """The same loss composed from torch ops, as kornia's ssim_loss + F.mse_loss launch it (baseline)."""
        x, y = out.moveaxis(-1, 0).unsqueeze(0), gt.moveaxis(-1, 0).unsqueeze(0)
        k = torch.exp(-(torch.arange(win, dtype=torch.float32, device=out.device) - win // 2).pow(2) / (2 * 1.5 ** 2))
        k = k / k.sum()

        def filt(img):
            z = F.pad(img, (win // 2,) * 4, mode="reflect")
            z = F.conv2d(z, k.view(1, 1, 1, -1).expand(3, 1, 1, -1), groups=3)
            return F.conv2d(z, k.view(1, 1, -1, 1).expand(3, 1, -1, 1), groups=3)

        mu1, mu2 = filt(x), filt(y)
        s1, s2, s12 = filt(x * x) - mu1 ** 2, filt(y * y) - mu2 ** 2, filt(x * y) - mu1 * mu2
        ssim = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 ** 2 + mu2 ** 2 + 1e-4) * (s1 + s2 + 9e-4) + 1e-12)
        return mult * torch.clamp((1 - ssim) / 2, 0, 1).mean() + (1 - mult) * F.mse_loss(out, gt)

    dev = torch.device("cuda:0")
    cam = S.make_camera("cfg2")
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(cam.h, cam.w, 3, generator=g).to(dev)
    out = (gt + 0.1 * torch.randn(cam.h, cam.w, 3, generator=g).to(dev)).requires_grad_(True)
    ref_fn = torch_loss_fn

    def ours():
        out.grad = None
        image_loss(out, gt, "l2", 0.2, 11).backward()

    def torch_ops():
        out.grad = None
        ref_fn(out, gt).backward()

    res = {}
    for name, fn in (("fused", ours), ("torch_ops", torch_ops)):
        for _ in range(3):
            fn()
        res[name] = timed(fn, 20, 1, dev)
    print(json.dumps({"config": "loss", "image": [cam.w, cam.h], "metric": "loss fwd+bwd ms (0.2 SSIM(11) + 0.8 L2)",
                      "ms": res, "speedup": res["torch_ops"] / res["fused"]}), flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    if which == "loss":
        return bench_loss()
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    scene_name = "cfg2" if which == "cfg4" else which
    cam = S.make_camera(scene_name)
    sc = S.make_scene(scene_name, seed=0)
    C = sc["C"]
    r = S.renderer_from_scene(sc, S.make_cfg(device=str(dev), sh_order=C))
    c2w = sc["c2w"].to(dev)
    out = {"config": which, "N": r.N, "C": C, "image": [cam.w, cam.h], "n_gpus": world, "steps": steps}
    if which == "cfg3":
        # the reference's loop body (main_sh.py:150-238): forward, loss, zero_grad, backward, step,
        # adaptive_control (statistics only between the structural iterations), optimiser RE-CREATED
        tgt = S.make_target(cam, 0).to(dev)
        variants = {}
        for vname, over in (("fused_single_step", dict(fused_adam=True, adam_single_step=True)),
                            ("fused", dict(fused_adam=True, adam_single_step=False)),
                            ("torch_adam", dict(fused_adam=False))):
            r = S.renderer_from_scene(sc, S.make_cfg(device=str(dev), sh_order=C, warm_up=0, **over))
            r.train()
            r.fuse_adc = True
            state = {"opt": r.get_optimizer(0), "e": 0}

            def step(r=r, state=state):
                o = r(c2w, cam)
                loss = ((o - tgt) ** 2).mean()
                state["opt"].zero_grad()
                loss.backward()
                state["opt"].step()
                r.adaptive_control(state["e"])
                state["opt"] = r.get_optimizer(state["e"])
                state["e"] += 1

            for _ in range(3):
                step()
            variants[vname] = timed(step, steps, world, dev)
            del r, state
            torch.cuda.empty_cache()
        ms = variants["fused_single_step"]
        out.update({"metric": "full training steps/s (fwd + L2 + bwd + ADC accumulation + Adam, optimiser "
                              "re-created every step like main_sh.py:238)",
                    "ms_per_step": ms, "value": 1000.0 / ms, "ms_per_step_by_optimiser": variants})
    elif which == "cfg4":
        mode = sys.argv[3] if len(sys.argv) > 3 else "pull"
        r.train()
        c2ws = [c.to(dev) for c in S.ring_cameras(8)]
        targets = [S.make_target(cam, i).to(dev) for i in range(8)]
        flat = P.FlatGradients(r, fused=(mode == "fused"), sparse=(mode == "sparse"), push=(mode == "push"),
                               pull=(mode == "pull"), sparse_reset=(world == 1)).attach(r)

        def step():
            P.view_sharded_step(r, flat, c2ws, cam, targets)

        for _ in range(3):
            step()
        ms = timed(step, steps, world, dev)
        out.update({"metric": "8-camera training steps/s (views sharded, gradients exchanged, ADC statistics synced)",
                    "exchange": mode if world > 1 else None, "ms_per_step": ms, "value": 1000.0 / ms,
                    "views_per_s": 8000.0 / ms})
    else:
        r.eval()

        def frame():
            return P.tile_sharded_render(r, c2w, cam)

        for _ in range(3):
            frame()
        ms = timed(frame, steps, world, dev)
        out.update({"metric": "forward FPS, tile rows sharded over ranks (all_gather of bands included)",
                    "ms_per_frame": ms, "value": 1000.0 / ms})
        if world == 1:
            out["n_dub"] = None
            with torch.no_grad():
                r(c2w, cam)
            out["n_dub"] = r.total_dub_gaussians
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
