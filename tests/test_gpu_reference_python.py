"""INTEGRATION.md route 2, executed: the reference's UNMODIFIED Python (`gs/sh_renderer.py`,
`gs/renderer.py`, `gs/culling.py`, `utils/camera.py`, ... staged byte-for-byte into the git-ignored
oracle/_ref/py by oracle/build_ref.py) runs with `import _gs` resolving to this repo's shim
(gaussian_splatting_3d_b200._gs -> libgs3d_b200.so), and is compared with route 1 (this repo's own
SHRenderer on the fused kernels) from the same leaf parameters:

  * SHRenderer.forward / backward (/root/reference/gs/sh_renderer.py:188-316) -- image, duplicate count and
    every leaf gradient;
  * a 5-iteration training loop shaped like /root/reference/main_sh.py:141-243 (forward, loss,
    opt.zero_grad, backward, opt.step, adaptive_control -> update_grads, optimiser re-created) -- per-iteration
    loss, the ADC accumulators (cnt, grad_mean) and the final image.

Third-party modules absent offline (torchtyping, kornia, matplotlib, faiss, cv2, ...) are stubbed exactly as
oracle/make_golden.py does.  Skipped when oracle/_ref/py did not travel to the box."""
import sys
from pathlib import Path

import pytest
import torch

from gaussian_splatting_3d_b200 import synthetic as S

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PY = ROOT / "oracle" / "_ref" / "py"
DEV = "cuda:0"
NAMES = ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")


@pytest.fixture(scope="module")
def refpy():
    if not (PY / "gs" / "sh_renderer.py").exists():
        pytest.skip("oracle/_ref/py (staged reference Python) not present")
    from oracle.make_golden import install_shims

    install_shims()
    import gaussian_splatting_3d_b200._gs as shim

    shim.install()  # sys.modules["_gs"] -> the reference's `import _gs as _backend` picks it up
    sys.path.insert(0, str(PY))
    try:
        import gs.sh_renderer as ref_sh  # noqa: E402  (the reference's module, unmodified)
        from utils.camera import CameraInfo  # noqa: E402
    finally:
        sys.path.remove(str(PY))
    assert ref_sh._backend.__name__ == "gaussian_splatting_3d_b200._gs"
    assert Path(ref_sh.__file__).resolve().is_relative_to(PY.resolve())
    return ref_sh, CameraInfo


def _cfg(C, **over):
    return S.make_cfg(device=DEV, sh_order=C, remove_low_alpha_period=0, eval_iteration=0, fused_adam=False, **over)


def _reference_renderer(ref_sh, sc, cfg):
    """The reference's SHRenderer holding the scene's parameters (its __init__ without pts leaves the
    parameters to load(); this is what SHRenderer.load does, sh_renderer.py:680-708)."""
    r = ref_sh.SHRenderer(cfg)
    r.N = sc["mean"].shape[0]
    for k in NAMES:
        setattr(r, k, torch.nn.Parameter(sc[k].to(DEV).clone()))
    r.register_buffer("grad_mean", torch.zeros(r.N, device=DEV))
    r.register_buffer("cnt", torch.zeros(r.N, dtype=torch.int32, device=DEV))
    r.now_C = sc["C"]
    return r


def _cam(CameraInfo, name):
    c = S.make_camera(name)
    return CameraInfo(c.fx, c.fy, c.cx, c.cy, c.w, c.h, c.near_plane, c.far_plane), c


def test_reference_forward_backward_on_shim_matches_fused_path(refpy):
    ref_sh, CameraInfo = refpy
    name, n, C = "cfg3", 200_000, 3
    sc = S.make_scene(name, seed=4, N=n, C=C)
    cam_ref, cam = _cam(CameraInfo, name)
    c2w = sc["c2w"].to(DEV)
    tgt = S.make_target(cam, 4).to(DEV)

    r2 = _reference_renderer(ref_sh, sc, _cfg(C))
    r2.train()
    out2 = r2(c2w, cam_ref)
    ((out2 - tgt) ** 2).mean().backward()

    r1 = S.renderer_from_scene(sc, _cfg(C))
    r1.train()
    out1 = r1(c2w, cam)
    ((out1 - tgt) ** 2).mean().backward()
    torch.cuda.synchronize()

    assert r1.total_dub_gaussians == r2.total_dub_gaussians
    err = (out1 - out2).abs()
    print(f"[route2] n_dub {r2.total_dub_gaussians}, image max-abs {float(err.max()):.3e}, "
          f"elements > 1e-4: {int((err > 1e-4).sum())}")
    assert float(err.max()) <= 1e-4
    for k in NAMES:
        a, b = getattr(r1, k).grad.double(), getattr(r2, k).grad.double()
        l2 = float((a - b).norm() / b.norm().clamp_min(1e-30))
        print(f"[route2] grad {k}: L2 rel {l2:.2e}")
        assert l2 <= 1e-3, (k, l2)
    # sh_renderer.py:215-221: cnt and the retained 2-D mean gradient
    assert torch.equal(r1.cnt, r2.cnt)
    g2 = torch.zeros(n, 2, device=DEV)
    g2[r2.frustum_culling_mask] = r2.mean_2d.grad
    d = (r1.mean_2d.grad.double() - g2.double()).norm() / g2.double().norm().clamp_min(1e-30)
    assert float(d) <= 1e-3


def test_reference_training_loop_on_shim_matches_fused_path(refpy):
    """main_sh.py:141-243 without the logging: the SAME loop body drives both renderers."""
    ref_sh, CameraInfo = refpy
    name, n, C = "cfg3", 100_000, 2
    sc = S.make_scene(name, seed=5, N=n, C=C)
    cam_ref, cam = _cam(CameraInfo, name)
    poses = [p.to(DEV) for p in S.ring_cameras(8, radius=2.0, centre=(0.0, 0.0, 4.0))][:3] + [sc["c2w"].to(DEV)]
    tgts = [S.make_target(cam, 10 + i).to(DEV) for i in range(len(poses))]
    over = dict(warm_up=0, split_type="2d_mean_grad", split_reduction="mean", max_iteration=10)

    def loop(renderer, camera_info, iters=5):
        losses = []
        renderer.train()
        opt = renderer.get_optimizer(0)
        for e in range(iters):
            i = e % len(poses)
            out = renderer(poses[i], camera_info)
            loss = ((out - tgts[i]) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            renderer.adaptive_control(e)
            opt = renderer.get_optimizer(e)
            losses.append(float(loss.item()))
        renderer.eval()
        with torch.no_grad():
            final = renderer(poses[-1], camera_info).clone()
        return losses, final

    r2 = _reference_renderer(ref_sh, sc, _cfg(C, **over))
    l2, img2 = loop(r2, cam_ref)
    r1 = S.renderer_from_scene(sc, _cfg(C, **over))
    l1, img1 = loop(r1, cam)
    print("[route2 loop] losses reference-python-on-shim:", [f"{x:.6f}" for x in l2])
    print("[route2 loop] losses fused path             :", [f"{x:.6f}" for x in l1])
    for a, b in zip(l1, l2):
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-12) + 1e-7, (l1, l2)
    assert torch.equal(r1.cnt, r2.cnt)
    gm = (r1.grad_mean.double() - r2.grad_mean.double()).norm() / r2.grad_mean.double().norm().clamp_min(1e-30)
    assert float(gm) <= 1e-3, float(gm)
    # after five Adam steps the two models are the same up to the gradients' atomic-order noise
    frac = float(((img1 - img2).abs() > 1e-3).float().mean())
    print(f"[route2 loop] final image: max-abs {float((img1 - img2).abs().max()):.3e}, elements > 1e-3: {frac:.2e}")
    assert frac <= 1e-3
