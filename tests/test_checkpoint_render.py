"""Checkpoint format + render-only entry: SHRenderer.save / load keep the reference's
on-disk format (gs/sh_renderer.py:663-708: a torch.save'd dict with the five tensors, N and cfg), and the
render entry reproduces the viewer / evaluation call `renderer(c2w, camera_info)` under no_grad."""
import numpy as np
import pytest
import torch

from gaussian_splatting_3d_b200 import synthetic as S

KEYS = {"mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation", "N", "cfg"}


def test_checkpoint_has_the_reference_format_and_round_trips_on_cpu(tmp_path):
    """Host logic only (no kernels): save writes exactly the reference's keys; a file written the
    reference's way (plain dict, sh_renderer.py:668-680) loads."""
    from gaussian_splatting_3d_b200.gs.sh_renderer import SHRenderer

    sc = S.make_scene("cfg1", seed=3, N=500)
    cfg = S.make_cfg(device="cpu", sh_order=sc["C"])
    r = S.renderer_from_scene(sc, cfg)
    r.save(tmp_path / "a" / "model.pt")
    state = torch.load(tmp_path / "a" / "model.pt", weights_only=False)
    assert set(state) == KEYS and state["N"] == 500
    ref_style = {k: sc[k].clone() for k in KEYS - {"N", "cfg"}}
    ref_style.update(N=500, cfg=cfg)
    torch.save(ref_style, tmp_path / "ref.pt")
    r2 = SHRenderer.load(tmp_path / "ref.pt")
    for k in KEYS - {"N", "cfg"}:
        assert torch.equal(getattr(r2, k).data, sc[k]), k
    assert r2.N == 500 and r2.grad_mean.shape == (500,) and r2.cnt.dtype == torch.int32


@pytest.mark.gpu
def test_render_entry_matches_the_model_forward_and_writes_frames(tmp_path):
    from gaussian_splatting_3d_b200 import render as RE

    dev = "cuda:0"
    cam = S.make_camera("cfg1")
    sc = S.make_scene("cfg1", seed=4)
    r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=sc["C"]))
    r.save(tmp_path / "m.pt")
    r.eval()
    poses = [sc["c2w"]] + S.ring_cameras(3, radius=7.0)
    with torch.no_grad():
        want = torch.stack([r(p.to(dev), cam) for p in poses])
    m = RE.load_model(tmp_path / "m.pt", dev, sh_order=sc["C"])
    frames, fps = RE.render_views(m, poses, cam, repeat=2)
    assert torch.equal(frames, want) and fps > 0  # same kernels, same inputs, deterministic forward
    np.save(tmp_path / "poses.npy", torch.stack(poses).numpy())
    rc = RE.main([str(tmp_path / "m.pt"), "--poses", str(tmp_path / "poses.npy"), "--camera", str(cam.fx),
                  str(cam.fy), str(cam.cx), str(cam.cy), str(cam.w), str(cam.h), "--near", str(cam.near_plane),
                  "--far", str(cam.far_plane), "--out-dir", str(tmp_path / "frames"), "--sh-order", str(sc["C"]),
                  "--device", dev])
    assert rc == 0
    got = np.load(tmp_path / "frames" / "frames.npy")
    assert np.array_equal(got, want.cpu().numpy())
    head = (tmp_path / "frames" / "frame_0000.ppm").read_bytes()[:15]
    assert head.startswith(f"P6\n{cam.w} {cam.h}\n255\n".encode())


def test_render_entry_host_logic_on_cpu(tmp_path):
    """No kernels: load_model restores a reference-style checkpoint (now_C from --sh-order or max_C), and
    write_ppm clamps like the reference's save_img and writes a valid binary PPM."""
    from gaussian_splatting_3d_b200 import render as RE

    sc = S.make_scene("cfg1", seed=8, N=64)
    cfg = S.make_cfg(device="cpu", sh_order=3)
    state = {k: sc[k].clone() for k in KEYS - {"N", "cfg"}}
    state["sh_coeffs"] = torch.zeros(64, 3, 9)
    state.update(N=64, cfg=cfg)
    torch.save(state, tmp_path / "m.pt")
    m = RE.load_model(tmp_path / "m.pt", "cpu")
    assert m.N == 64 and m.now_C == 3 and not m.training
    assert RE.load_model(tmp_path / "m.pt", "cpu", sh_order=2).now_C == 2
    img = torch.tensor([[[-0.5, 0.5, 2.0], [0.0, 1.0, 0.25]]])  # [1, 2, 3]
    RE.write_ppm(tmp_path / "a.ppm", img)
    raw = (tmp_path / "a.ppm").read_bytes()
    assert raw.startswith(b"P6\n2 1\n255\n")
    assert list(raw[-6:]) == [0, 128, 255, 0, 255, 64]


def test_fused_adam_argument_validation():
    from gaussian_splatting_3d_b200.optim import FusedAdam

    w = torch.nn.Parameter(torch.zeros(3))
    with pytest.raises(ValueError):
        FusedAdam([w], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([w], betas=(1.0, 0.99))
    opt = FusedAdam([{"params": [w], "lr": 0.5}], lr=1e-3, betas=(0.9, 0.99), single_step=True)
    assert opt.param_groups[0]["lr"] == 0.5 and opt.param_groups[0]["betas"] == (0.9, 0.99)
    assert opt.step() is None  # no gradients: nothing to launch, no CUDA needed
    opt.zero_grad()
