"""CPU tests of the host-side mirror of the reference interface (no GPU)."""
import numpy as np
import torch

from gaussian_splatting_3d_b200 import synthetic as S
from gaussian_splatting_3d_b200.utils.camera import CameraInfo, get_c2w_from_up_and_look_at, in_frustum
from gaussian_splatting_3d_b200.utils.transforms import (qsvec2rotmat_batched, qvec2rotmat, qvec2rotmat_batched,
                                                         rotmat2wxyz)
from oracle import ref_torch as R


def test_camera_info_matches_reference_semantics():
    cam = CameraInfo(961.22, 963.09, 648.38, 420.12, 1297, 840, 0.0, 1000)
    assert abs(cam.yfov - 2 * np.arctan(840 / (2 * 963.09))) < 1e-12 and cam.aspect == 1297 / 840
    cam.upsample(4)
    assert (cam.w, cam.h) == (5188, 3360) and abs(cam.fx - 4 * 961.22) < 1e-9
    cam.downsample(4)
    assert (cam.w, cam.h) == (1297, 840)
    f = CameraInfo.from_fov_camera(np.pi / 3, 16 / 9, 1280, 0.1, 100.0)
    assert f.w == 1280 and f.h == 720 and abs(f.fx - 640 / np.tan(np.pi / 6)) < 1e-9
    px = cam.camera_space_to_pixel_space(torch.tensor([[0.1, -0.2], [-1.0, 0.5]]))
    assert px.dtype == torch.int32
    assert px.tolist() == [[int(0.1 * 961.22 + 648.38), int(np.float32(-0.2) * np.float32(963.09) + np.float32(420.12))],
                           [int(-961.22 + 648.38), int(0.5 * 963.09 + 420.12)]]


def test_cpu_frustum_matches_oracle_and_pose_helper():
    c2w = torch.from_numpy(get_c2w_from_up_and_look_at(np.array([0, 0, 1.0]), np.zeros(3), np.array([1.0, 0, 0])))
    assert c2w.shape == (3, 4) and c2w.dtype == torch.float32
    np.testing.assert_allclose(c2w[:, 2].numpy(), [-1, 0, 0], atol=1e-7)  # looks at the origin
    cam = CameraInfo(300.0, 300.0, 128.0, 128.0, 256, 256, 0.5, 100.0)
    n1, p1 = cam.get_frustum(c2w)
    n2, p2 = R.get_frustum(c2w, cam)
    assert torch.equal(n1, n2) and torch.equal(p1, p2)
    inside = in_frustum(torch.tensor([[0.0, 0.0, 0.0], [5.0, 0.0, 0.0], [0.0, 50.0, 0.0]]), n1, p1)
    assert inside.tolist() == [True, False, False]


def test_quaternion_helpers():
    g = torch.Generator().manual_seed(0)
    q = torch.randn(16, 4, generator=g)
    Rm = qvec2rotmat_batched(q)
    np.testing.assert_allclose((Rm @ Rm.transpose(-1, -2)).numpy(), np.tile(np.eye(3), (16, 1, 1)), atol=1e-5)
    qn = (q / q.norm(dim=-1, keepdim=True)).double().numpy()
    for i in range(16):
        np.testing.assert_allclose(Rm[i].numpy(), qvec2rotmat(qn[i]), atol=1e-6)
    s = torch.rand(16, 3, generator=g)
    np.testing.assert_allclose(qsvec2rotmat_batched(q, s).numpy(), (Rm * s[:, None, :]).numpy())
    back = rotmat2wxyz(Rm)
    sign = torch.sign((back * q).sum(-1, keepdim=True))
    np.testing.assert_allclose((back * sign).numpy(), (q / q.norm(dim=-1, keepdim=True)).numpy(), atol=1e-5)


def test_synthetic_scene_is_deterministic_and_in_recipe():
    a = S.make_scene("cfg1", seed=0)
    b = S.make_scene("cfg1", seed=0)
    for k in ("mean", "qvec", "svec_before_activation", "alpha_before_activation", "sh_coeffs"):
        assert torch.equal(a[k], b[k])
    assert a["mean"].shape == (10_000, 3) and a["sh_coeffs"].shape == (10_000, 3, 1)
    z = a["mean"][:, 2]
    assert float(z.min()) >= 2.0 and float(z.max()) <= 12.0
    s = a["svec_before_activation"].exp()
    assert float(s.min()) >= 0.004 * 0.999 and float(s.max()) <= 0.04 * 1.001
    assert not torch.equal(a["mean"], S.make_scene("cfg1", seed=1)["mean"])
    cams = S.ring_cameras(8)
    assert len(cams) == 8
    for c in cams:
        Rm = c[:, :3]
        np.testing.assert_allclose((Rm.T @ Rm).numpy(), np.eye(3), atol=1e-5)
        np.testing.assert_allclose(float(torch.det(Rm)), 1.0, atol=1e-5)
        centre = torch.tensor([0.0, 0.0, 7.0])
        to_c = (centre - c[:, 3]) / (centre - c[:, 3]).norm()
        np.testing.assert_allclose(c[:, 2].numpy(), to_c.numpy(), atol=1e-5)


def test_schedulers_and_activations():
    from gaussian_splatting_3d_b200.utils.activations import activations, inv_activations
    from gaussian_splatting_3d_b200.utils.schedulers import lr_schedulers

    assert lr_schedulers["nothing"](100, 1e-3, 1e-5)(50) == 1e-3
    f = lr_schedulers["exp"](100, 1e-2, 1e-4, 10)
    assert f(0) == 0 and abs(f(10) - 1e-2) < 1e-12 and abs(f(100) - 1e-4) < 1e-12
    c = lr_schedulers["cosine"](100, 1e-2, 1e-4, 0)
    assert abs(c(0) - 1e-2) < 1e-12 and abs(c(100) - 1e-4) < 1e-12
    assert abs(activations["sigmoid"](torch.tensor(inv_activations["sigmoid"](0.3))).item() - 0.3) < 1e-6
    assert abs(inv_activations["exp"](0.01) - np.log(0.01)) < 1e-12
    x = torch.tensor([0.2, 0.7])
    assert torch.allclose(activations["sigmoid"](inv_activations["sigmoid"](x)), x, atol=1e-6)


def test_bench_weak_scaling_poses_are_centred_and_small():
    """bench.views_for: one pose per rank, yawed `yaw_step` degrees apart around the cfg-2 pose; rank poses of the
    default step stay inside the scene's 5 % margin (about +-3.8 degrees), world == 1 is the identity pose."""
    import math
    import sys
    from pathlib import Path

    import torch

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench

    assert torch.equal(bench.views_for(1, torch)[0], torch.eye(3, 4))
    for world in (2, 4, 8):
        views = bench.views_for(world, torch, 0.5)
        assert len(views) == world
        yaws = [math.degrees(math.atan2(float(v[0, 2]), float(v[0, 0]))) for v in views]
        assert abs(sum(yaws)) < 1e-4 and max(abs(y) for y in yaws) <= 0.5 * (world - 1) / 2 + 1e-4 < 3.8
        for a, b in zip(yaws[:-1], yaws[1:]):
            assert abs((b - a) - 0.5) < 1e-4
        for v in views:  # proper rotations, no translation
            R = v[:, :3]
            assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-6) and float(v[:, 3].abs().max()) == 0.0
