"""Render-only entry: the reference ships an EMPTY `render.py`; its render-only
callers are the evaluation loop (main_sh.py:206-211) and the viewer hook (utils/viewer/viser_viewer.py:
119-139), both of which just call `renderer(c2w, camera_info)` under `torch.no_grad()` on a model restored
with `SHRenderer.load` (gs/sh_renderer.py:682-708).  This module is that flow as a function and a CLI:

    python -m gaussian_splatting_3d_b200.render CKPT.pt --poses poses.npy --camera fx fy cx cy W H \
        [--near 0.1 --far 1000] [--out-dir frames/] [--sh-order C] [--repeat K]

`poses.npy` holds c2w matrices [V,3,4] (OpenCV axes).  Frames are written as binary PPM (no image library
needed) and as one float32 `frames.npy`; the forward FPS is measured with CUDA events over the loop.
CUDA only: there is no CPU fallback.
"""
import argparse
import json
from pathlib import Path

import numpy as np
import torch


def load_model(path, device="cuda", cfg=None, sh_order=None):
    """SHRenderer.load on `device`, in eval mode.  `sh_order` sets now_C (the checkpoint format does not
    store it: sh_renderer.py:663-680 saves the five tensors, N and cfg only); default max_C."""
    from .gs.sh_renderer import SHRenderer

    state = torch.load(path, map_location=device, weights_only=False)
    use_cfg = cfg if cfg is not None else state["cfg"]
    try:
        use_cfg.device = device
    except Exception:
        pass
    r = SHRenderer(use_cfg)
    r._set_params({k: state[k].to(device).contiguous() for k in
                   ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")})
    assert r.N == state["N"], "checkpoint N does not match its tensors"
    r._reset_adc_buffers()
    r.now_C = int(sh_order) if sh_order else r.max_C
    r.eval()
    return r


@torch.no_grad()
def render_views(renderer, c2ws, camera_info, repeat=1):
    """-> (frames [V,H,W,3] float32 on the device, forward FPS).  The timed loop is exactly the viewer's /
    evaluation loop's call: renderer(c2w, camera_info)."""
    dev = renderer.mean.device
    c2ws = [torch.as_tensor(c, dtype=torch.float32, device=dev).contiguous() for c in c2ws]
    frames = torch.empty(len(c2ws), camera_info.h, camera_info.w, 3, dtype=torch.float32, device=dev)
    for c in c2ws[:1]:
        renderer(c, camera_info)  # warm-up (module load, allocator)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(max(1, int(repeat))):
        for i, c in enumerate(c2ws):
            frames[i] = renderer(c, camera_info)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    fps = 1000.0 * len(c2ws) * max(1, int(repeat)) / ms if ms > 0 else float("inf")
    return frames, fps


def write_ppm(path, img):
    """img [H,W,3] float in [0,1] -> binary PPM (the reference's save_img clamps the same way)."""
    a = (img.clamp(0.0, 1.0) * 255.0).round().to(torch.uint8).cpu().numpy()
    with open(path, "wb") as f:
        f.write(f"P6\n{a.shape[1]} {a.shape[0]}\n255\n".encode())
        f.write(a.tobytes())


def main(argv=None):
    from .utils.camera import CameraInfo

    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("checkpoint")
    ap.add_argument("--poses", required=True, help=".npy with c2w [V,3,4]")
    ap.add_argument("--camera", nargs=6, type=float, required=True, metavar=("fx", "fy", "cx", "cy", "W", "H"))
    ap.add_argument("--near", type=float, default=0.1)
    ap.add_argument("--far", type=float, default=1000.0)
    ap.add_argument("--out-dir", default=None)
    ap.add_argument("--sh-order", type=int, default=None)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args(argv)
    fx, fy, cx, cy, W, H = a.camera
    cam = CameraInfo(fx, fy, cx, cy, int(W), int(H), a.near, a.far)
    poses = np.load(a.poses).astype(np.float32).reshape(-1, 3, 4)
    r = load_model(a.checkpoint, a.device, sh_order=a.sh_order)
    frames, fps = render_views(r, list(poses), cam, a.repeat)
    if a.out_dir:
        out = Path(a.out_dir)
        out.mkdir(parents=True, exist_ok=True)
        for i in range(frames.shape[0]):
            write_ppm(out / f"frame_{i:04d}.ppm", frames[i])
        np.save(out / "frames.npy", frames.cpu().numpy())
    print(json.dumps({"views": int(frames.shape[0]), "N": r.N, "C": r.now_C, "image": [int(W), int(H)],
                      "forward_fps": fps, "n_dub_last": r.total_dub_gaussians}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
