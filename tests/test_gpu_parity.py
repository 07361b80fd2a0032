"""GPU parity tests (run on the B200 with `-m gpu`): every kernel is called through the C ABI
(gaussian_splatting_3d_b200.ops -> capi -> libgs3d_b200.so) and compared with the CPU oracle on
the same seeded inputs, against the committed golden fixtures of the real reference, and -- at
BASELINE sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): tile keys / sort order / ranges bit-exact; images <= 1e-4
max-abs; gradients <= 1e-3 relative.  A skip decision (alpha*G vs 1/255) that flips because of a
1-ulp difference moves a pixel by up to 1/255, so image comparisons against the CPU oracle (whose
expf is glibc's, not CUDA's) are made on the pixels whose decisions are FP-stable (oracle margin
diagnostic) and the number of fragile pixels is bounded separately.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from gaussian_splatting_3d_b200 import synthetic as S

ROOT = Path(__file__).resolve().parent.parent

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from gaussian_splatting_3d_b200 import ops as o

    return o


@pytest.fixture(scope="module")
def K():
    from oracle import gs_oracle

    return gs_oracle


@pytest.fixture(scope="module")
def R():
    from oracle import ref_torch

    return ref_torch


def _cam_from(arr):
    from gaussian_splatting_3d_b200.utils.camera import CameraInfo

    fx, fy, cx, cy, w, h, near, far = arr.tolist()
    return CameraInfo(fx, fy, cx, cy, int(w), int(h), near, far)


def _rel(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


GOLDEN = ["project_cfg1_identity", "project_cfg1_posed", "project_cfg2_identity", "project_cfg2_posed",
          "project_cfg3_identity", "project_cfg3_posed"]


# ---------------------------------------------------------------- a1 / a2

@pytest.mark.parametrize("name", GOLDEN[:4])
def test_frustum_and_cull(ops, K, golden_dir, name):
    z = np.load(golden_dir / f"{name}.npz")
    cam = _cam_from(z["cam"])
    c2w = torch.from_numpy(z["c2w"]).to(DEV)
    normals, pts = ops.get_frustum(c2w, cam)
    np.testing.assert_allclose(normals.cpu().numpy(), z["f_normals"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(pts.cpu().numpy(), z["f_pts"], rtol=2e-6, atol=2e-7)
    # cull with the reference's own planes: mask must equal the oracle except for spheres within
    # an ulp of a plane (none expected at these sizes)
    mean, svec = z["mean"], z["svec"]
    mask = torch.zeros(mean.shape[0], dtype=torch.bool, device=DEV)
    ops.culling_gaussian_bsphere(torch.from_numpy(mean).to(DEV), torch.from_numpy(z["qvec"]).to(DEV),
                                 torch.from_numpy(svec).to(DEV), torch.from_numpy(z["f_normals"]).to(DEV),
                                 torch.from_numpy(z["f_pts"]).to(DEV), mask, 1.0)
    want = K.culling_gaussian_bsphere(mean, svec, z["f_normals"], z["f_pts"], 1.0)
    assert int((mask.cpu().numpy() != want).sum()) == 0
    assert 0 < want.sum() < want.size or name.endswith("identity")


# ---------------------------------------------------------------- a4 / a5

@pytest.mark.parametrize("name", GOLDEN)
def test_project_gaussians_vs_real_reference(ops, golden_dir, name):
    from gaussian_splatting_3d_b200.gs.renderer import project_gaussians

    z = np.load(golden_dir / f"{name}.npz")
    mean = torch.from_numpy(z["mean"]).to(DEV).requires_grad_(True)
    qvec = torch.from_numpy(z["qvec"]).to(DEV).requires_grad_(True)
    svec = torch.from_numpy(z["svec"]).to(DEV).requires_grad_(True)
    c2w = torch.from_numpy(z["c2w"]).to(DEV)
    mean2d, cov, JW, depth = project_gaussians(mean, qvec, svec, c2w, True)
    assert _rel(mean2d.detach().cpu(), z["mean2d"]) < 1e-5
    np.testing.assert_allclose(cov.detach().cpu().numpy(), z["cov"], rtol=2e-4, atol=1e-9)
    assert _rel(cov.detach().cpu(), z["cov"]) < 1e-5
    assert _rel(JW.detach().cpu(), z["JW"]) < 1e-5
    assert _rel(depth.detach().cpu(), z["depth"]) < 1e-6
    up_m = torch.from_numpy(z["up_mean2d"]).to(DEV)
    up_c = torch.from_numpy(z["up_cov"]).to(DEV)
    ((mean2d * up_m).sum() + (cov * up_c).sum()).backward()
    assert _rel(mean.grad.cpu(), z["g_mean"]) < 1e-4
    assert _rel(qvec.grad.cpu(), z["g_qvec"]) < 1e-4
    assert _rel(svec.grad.cpu(), z["g_svec"]) < 1e-4


@pytest.mark.parametrize("name", GOLDEN)
def test_tile_rects_bit_exact(golden_dir, name):
    from gaussian_splatting_3d_b200.gs.culling import tile_culling_aabb_count

    z = np.load(golden_dir / f"{name}.npz")
    cam = _cam_from(z["cam"])
    n, tl, br = tile_culling_aabb_count(torch.from_numpy(z["mean2d"]).to(DEV), torch.from_numpy(z["cov"]).to(DEV),
                                        16, cam, 6.0)
    assert n == int(z["n_dub"])
    assert np.array_equal(tl.cpu().numpy(), z["tl"])
    assert np.array_equal(br.cpu().numpy(), z["br"])


# ---------------------------------------------------------------- a6 binning

def _oracle_aux(R, name, seed, n, C=None, **kw):
    sc = S.make_scene(name, seed=seed, N=n, C=C)
    p = {k: sc[k] for k in ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")}
    img, aux = R.reference_forward(p, sc["c2w"], sc["camera"], sc["C"], return_aux=True, **kw)
    return sc, img, aux


@pytest.mark.parametrize("name,seed,n", [("cfg1", 0, 10_000), ("cfg2", 1, 40_000), ("cfg3", 2, 1), ("cfg1", 3, 37)])
def test_binning_bit_exact(ops, R, name, seed, n):
    sc, _, aux = _oracle_aux(R, name, seed, n)
    cam = sc["camera"]
    nth, ntw = (cam.h + 15) // 16, (cam.w + 15) // 16
    tl = aux["tl"].to(DEV).contiguous()
    br = aux["br"].to(DEV).contiguous()
    depth = aux["depth"].detach().to(DEV).contiguous()
    n_dub = aux["n_dub"]
    ids = torch.zeros(n_dub, dtype=torch.int32, device=DEV)
    start = torch.zeros(nth * ntw, dtype=torch.int32, device=DEV)
    end = torch.zeros(nth * ntw, dtype=torch.int32, device=DEV)
    keys = torch.zeros(n_dub, dtype=torch.int64, device=DEV)
    ops.tile_culling_aabb_start_end(tl, br, ids, start, end, depth, nth, ntw, sorted_keys=keys)
    assert np.array_equal(keys.cpu().numpy(), aux["keys"]), "sorted int64 keys differ"
    assert np.array_equal(start.cpu().numpy(), aux["start"])
    assert np.array_equal(end.cpu().numpy(), aux["end"])
    got, want = ids.cpu().numpy(), aux["ids"]
    if not np.array_equal(got, want):
        # ties (same tile, same depth bits) may legally differ: compare as multisets per key run
        k = aux["keys"]
        order_g = np.lexsort((got, k))
        order_w = np.lexsort((want, k))
        assert np.array_equal(got[order_g], want[order_w])
    # idempotence (gs/debug.py:991-1023)
    ids2 = torch.zeros_like(ids)
    s2, e2 = torch.zeros_like(start), torch.zeros_like(end)
    ops.tile_culling_aabb_start_end(tl, br, ids2, s2, e2, depth, nth, ntw)
    assert torch.equal(ids2, ids) and torch.equal(s2, start) and torch.equal(e2, end)


def test_binning_count_mismatch_raises(ops):
    tl = torch.zeros(4, 2, dtype=torch.int32, device=DEV)
    br = torch.ones(4, 2, dtype=torch.int32, device=DEV)
    depth = torch.rand(4, 1, device=DEV)
    ids = torch.zeros(15, dtype=torch.int32, device=DEV)  # rects add up to 16
    se = torch.zeros(4, dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError):
        ops.tile_culling_aabb_start_end(tl, br, ids, se, se.clone(), depth, 2, 2)


def test_binning_negative_depth_and_ties(ops, K):
    # negative depths sort AFTER positive ones (raw bit order, quirk Q9); equal depths tie by id
    n = 3000
    g = torch.Generator().manual_seed(5)
    depth = torch.randn(n, 1, generator=g)
    depth[::7] = 1.25  # many exact ties
    tlx = torch.randint(0, 6, (n,), generator=g)
    tly = torch.randint(0, 5, (n,), generator=g)
    w = torch.randint(0, 3, (n,), generator=g)
    h = torch.randint(0, 4, (n,), generator=g)
    tl = torch.stack([tlx, tly], 1).int()
    br = torch.stack([torch.minimum(tlx + w, torch.tensor(7)), torch.minimum(tly + h, torch.tensor(7))], 1).int()
    n_dub = int(((br[:, 0] - tl[:, 0] + 1) * (br[:, 1] - tl[:, 1] + 1)).sum())
    want_ids, want_s, want_e, want_k = K.tile_culling_aabb_start_end(tl.numpy(), br.numpy(), depth.numpy(), n_dub, 8, 8)
    ids = torch.zeros(n_dub, dtype=torch.int32, device=DEV)
    start = torch.zeros(64, dtype=torch.int32, device=DEV)
    end = torch.zeros(64, dtype=torch.int32, device=DEV)
    keys = torch.zeros(n_dub, dtype=torch.int64, device=DEV)
    ops.tile_culling_aabb_start_end(tl.to(DEV), br.to(DEV), ids, start, end, depth.to(DEV), 8, 8, sorted_keys=keys)
    assert np.array_equal(keys.cpu().numpy(), want_k)
    assert np.array_equal(ids.cpu().numpy(), want_ids)
    assert np.array_equal(start.cpu().numpy(), want_s) and np.array_equal(end.cpu().numpy(), want_e)


# ---------------------------------------------------------------- a7 / a8 compositing

def _render_inputs(aux, sc, C):
    cam = sc["camera"]
    H, W = cam.h, cam.w
    nth, ntw = (H + 15) // 16, (W + 15) // 16
    t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray)  # noqa: E731
                            else a.detach().contiguous()).to(DEV)
    d = dict(
        mean=t(aux["mean2d"]), cov=t(aux["cov"]), sh=t(aux["sh"][..., : C * C].detach().contiguous()),
        alpha=t(aux["alpha"]), start=t(aux["start"]), end=t(aux["end"]), ids=t(aux["ids"]),
        topleft=torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32, device=DEV),
        c2w=sc["c2w"].to(DEV).contiguous(),
        consts=(16, nth, ntw, 1.0 / cam.fx, 1.0 / cam.fy, H, W, C, 1e-4),
    )
    return d


def _check_image(got, want, margin, tag, stable_thr=2e-3, max_fragile_frac=0.02):
    got, want = got.reshape(-1, 3), want.reshape(-1, 3)
    err = np.abs(got - want).max(axis=1)
    stable = margin > stable_thr
    assert stable.mean() > 1 - max_fragile_frac, f"{tag}: too many FP-fragile pixels ({1 - stable.mean():.4f})"
    worst = err[stable].max() if stable.any() else 0.0
    assert worst <= 1e-4, f"{tag}: max-abs {worst:.3e} on decision-stable pixels (n={stable.sum()})"
    # fragile pixels can only be off by one skipped/added splat: <= 1/255 * (1 + small)
    assert err.max() <= 2.5 / 255, f"{tag}: fragile-pixel error {err.max():.3e} exceeds two flipped splats"
    return worst, float(1 - stable.mean()), float(err.max())


@pytest.mark.parametrize("name,seed,n,C", [("cfg1", 0, 10_000, 1), ("cfg1", 1, 6_000, 2), ("cfg3", 2, 30_000, 3),
                                           ("cfg2", 3, 60_000, 4)])
@pytest.mark.parametrize("exact", [True, False])
def test_render_forward_vs_oracle(K, R, name, seed, n, C, exact):
    import gaussian_splatting_3d_b200._gs as gs

    sc, img, aux = _oracle_aux(R, name, seed, n, C=C)
    d = _render_inputs(aux, sc, C)
    consts = d["consts"]
    cam = sc["camera"]
    _, fT, nc, margin = K.render_sh_forward(aux["mean2d"].detach().numpy(), aux["cov"].detach().numpy(),
                                            aux["sh"][..., : C * C].detach().contiguous().numpy(),
                                            aux["alpha"].detach().numpy(), aux["start"], aux["end"], aux["ids"],
                                            d["topleft"].cpu().numpy(), sc["c2w"].numpy(), consts[0], consts[1],
                                            consts[2], np.float32(consts[3]), np.float32(consts[4]), *consts[5:],
                                            diagnostics=True)
    out = torch.zeros(cam.h * cam.w * 3, device=DEV)
    gs.EXACT_DECISIONS = exact
    try:
        gs.tile_based_vol_rendering_sh(d["mean"], d["cov"], d["sh"], d["alpha"], d["start"], d["end"], d["ids"], out,
                                       d["topleft"], d["c2w"], *consts)
    finally:
        gs.EXACT_DECISIONS = True
    worst, frag, emax = _check_image(out.cpu().numpy(), img.detach().numpy(), margin, f"{name}/C{C}/exact={exact}")
    print(f"[fwd {name} C={C} exact={exact}] stable max-abs {worst:.2e} fragile frac {frag:.4f} overall max {emax:.2e}")


@pytest.mark.parametrize("name,seed,n,C", [("cfg1", 0, 10_000, 1), ("cfg1", 1, 6_000, 2), ("cfg3", 2, 30_000, 3),
                                           ("cfg2", 3, 60_000, 4)])
def test_render_backward_vs_oracle(K, R, name, seed, n, C):
    import gaussian_splatting_3d_b200._gs as gs

    sc, img, aux = _oracle_aux(R, name, seed, n, C=C)
    d = _render_inputs(aux, sc, C)
    consts = d["consts"]
    cam = sc["camera"]
    tgt = S.make_target(cam, seed)
    g_out = (2.0 * (img.detach() - tgt) / img.numel()).reshape(-1).contiguous()
    out_o = img.detach().reshape(-1).numpy()
    want = K.render_sh_backward(aux["mean2d"].detach().numpy(), aux["cov"].detach().numpy(),
                                aux["sh"][..., : C * C].detach().contiguous().numpy(), aux["alpha"].detach().numpy(),
                                aux["start"], aux["end"], aux["ids"], out_o, g_out.numpy(),
                                d["topleft"].cpu().numpy(), sc["c2w"].numpy(), consts[0], consts[1], consts[2],
                                np.float32(consts[3]), np.float32(consts[4]), *consts[5:])
    out = torch.zeros(cam.h * cam.w * 3, device=DEV)
    gs.tile_based_vol_rendering_sh(d["mean"], d["cov"], d["sh"], d["alpha"], d["start"], d["end"], d["ids"], out,
                                   d["topleft"], d["c2w"], *consts)
    gm = torch.zeros_like(d["mean"])
    gc = torch.zeros_like(d["cov"])
    gsh = torch.zeros_like(d["sh"])
    ga = torch.zeros_like(d["alpha"])
    gs.tile_based_vol_rendering_backward_sh(d["mean"], d["cov"], d["sh"], d["alpha"], d["start"], d["end"], d["ids"],
                                            out, gm, gc, gsh, ga, g_out.to(DEV), d["topleft"], d["c2w"], *consts)
    for tag, got, w in (("mean", gm, want[0]), ("cov", gc.reshape(-1, 4), want[1]), ("sh", gsh, want[2]),
                        ("alpha", ga, want[3])):
        r = _rel(got.cpu().numpy(), w)
        print(f"[bwd {name} C={C}] grad_{tag}: rel max err {r:.2e} (|want|max {np.abs(w).max():.3e})")
        # Against the CPU oracle a handful of 1/255 skip decisions differ (glibc expf vs CUDA expf:
        # each flip adds or removes one splat's whole contribution to one pixel), so the bound here
        # is 3e-3 in the L2 sense; the 1e-3 bound of the north star is enforced like-for-like against
        # the real reference extension in tests/test_gpu_vs_reference_ext.py.
        l2 = np.linalg.norm(got.cpu().numpy().astype(np.float64).ravel() - w.astype(np.float64).ravel()) / \
            max(np.linalg.norm(w.astype(np.float64).ravel()), 1e-30)
        print(f"[bwd {name} C={C}] grad_{tag}: L2 rel {l2:.2e}")
        assert l2 <= 3e-3, f"grad_{tag} L2 rel err {l2:.3e}"
        assert r <= 2e-2, f"grad_{tag} max rel err {r:.3e}"


def test_render_bg_and_empty_tiles(K, R):
    import gaussian_splatting_3d_b200._gs as gs

    # few Gaussians on a 3x3-tile image with ragged edges: most tiles empty
    from gaussian_splatting_3d_b200.utils.camera import CameraInfo

    cam = CameraInfo(40.0, 40.0, 19.0, 21.0, 41, 37, 0.5, 100.0)
    sc = S.make_scene(None, seed=4, N=12, C=2, camera=cam)
    sc["svec_before_activation"] += 1.0
    p = {k: sc[k] for k in ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")}
    bg = [0.25, 0.5, 0.75]
    img, aux = R.reference_forward(p, sc["c2w"], cam, 2, return_aux=True, bg_rgb=bg)
    d = _render_inputs(aux, sc, 2)
    out = torch.zeros(cam.h * cam.w * 3, device=DEV)
    gs.tile_based_vol_rendering_sh_with_bg(d["mean"], d["cov"], d["sh"], d["alpha"], d["start"], d["end"], d["ids"],
                                           out, d["topleft"], d["c2w"], *d["consts"],
                                           torch.tensor(bg, device=DEV))
    err = np.abs(out.cpu().numpy() - img.detach().reshape(-1).numpy()).max()
    assert err <= 1e-4, f"bg variant max-abs {err:.3e}"
    assert (aux["start"] == -1).any()


def test_render_rejects_bad_arguments(ops):
    import gaussian_splatting_3d_b200._gs as gs

    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)  # noqa: E731
    args = [z(1, 2), z(1, 2, 2), z(1, 3, 1), z(1), z(1, dt=torch.int32), z(1, dt=torch.int32), z(1, dt=torch.int32),
            z(16 * 16 * 3), z(2), z(3, 4)]
    with pytest.raises(RuntimeError):  # tile_size != 16
        gs.tile_based_vol_rendering_sh(*args, 8, 1, 1, 0.1, 0.1, 16, 16, 1, 1e-4)
    with pytest.raises(RuntimeError):  # C out of range
        gs.tile_based_vol_rendering_sh(*args, 16, 1, 1, 0.1, 0.1, 16, 16, 5, 1e-4)
    bad = list(args)
    bad[0] = bad[0].cpu()
    with pytest.raises(RuntimeError):  # CPU tensor
        gs.tile_based_vol_rendering_sh(*bad, 16, 1, 1, 0.1, 0.1, 16, 16, 1, 1e-4)
    bad = list(args)
    bad[4] = bad[4].float()
    with pytest.raises(RuntimeError):  # wrong dtype
        gs.tile_based_vol_rendering_sh(*bad, 16, 1, 1, 0.1, 0.1, 16, 16, 1, 1e-4)


# ---------------------------------------------------------------- whole path through SHRenderer

@pytest.mark.parametrize("name,seed,n,C,maxC", [("cfg1", 0, 10_000, 1, 1), ("cfg3", 1, 20_000, 3, 3),
                                                ("cfg2", 2, 50_000, 4, 4), ("cfg1", 3, 5_000, 2, 4)])
def test_shrenderer_forward_backward_vs_reference_flow(K, R, name, seed, n, C, maxC):
    """SHRenderer (fused path, no compaction) vs the reference's flow restated on the CPU
    (sh_renderer.py:188-316): image, leaf gradients, ADC buffers."""
    cam = S.make_camera(name)
    sc = S.make_scene(name, seed=seed, N=n, C=C, max_C=maxC)
    cfg = S.make_cfg(device=DEV, sh_order=maxC)
    r = S.renderer_from_scene(sc, cfg)
    r.train()
    tgt = S.make_target(cam, seed)
    out = r(sc["c2w"].to(DEV), cam)
    loss = ((out - tgt.to(DEV)) ** 2).mean()
    loss.backward()

    names = ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")
    p = {k: sc[k].clone().requires_grad_(True) for k in names}
    img, aux = R.reference_forward(p, sc["c2w"], cam, C, return_aux=True)
    ((img - tgt) ** 2).mean().backward()

    assert r.total_dub_gaussians == aux["n_dub"], (r.total_dub_gaussians, aux["n_dub"])
    mask = r.frustum_culling_mask.cpu().numpy()
    assert int((mask != aux["mask"].numpy()).sum()) == 0
    # margin diagnostic from the oracle on its own projected inputs
    consts = (16, (cam.h + 15) // 16, (cam.w + 15) // 16, np.float32(1 / cam.fx), np.float32(1 / cam.fy), cam.h,
              cam.w, C, 1e-4)
    _, _, _, margin = K.render_sh_forward(aux["mean2d"].detach().numpy(), aux["cov"].detach().numpy(),
                                          aux["sh"][..., : C * C].detach().contiguous().numpy(),
                                          aux["alpha"].detach().numpy(), aux["start"], aux["end"], aux["ids"],
                                          np.array([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=np.float32),
                                          sc["c2w"].numpy(), *consts, diagnostics=True)
    worst, frag, emax = _check_image(out.detach().cpu().numpy(), img.detach().numpy(), margin, f"e2e {name}",
                                     stable_thr=5e-3, max_fragile_frac=0.05)
    print(f"[e2e {name} C={C}] image stable max-abs {worst:.2e}, fragile {frag:.4f}, overall {emax:.2e}")
    for k in names:
        got = getattr(r, k).grad.cpu().numpy().astype(np.float64)
        want = p[k].grad.numpy().astype(np.float64)
        l2 = np.linalg.norm((got - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-30)
        print(f"[e2e {name} C={C}] grad {k}: L2 rel {l2:.2e}")
        assert l2 <= 2e-3, f"{k}: L2 rel err {l2:.3e}"
    # ADC bookkeeping (sh_renderer.py:215-221, 602-623)
    assert np.array_equal(r.cnt.cpu().numpy(), aux["mask"].numpy().astype(np.int32))
    r.update_grads()
    want_gm = np.zeros(n, dtype=np.float32)
    want_gm[aux["mask"].numpy()] = aux["mean2d"].grad.norm(dim=-1).numpy()
    gm = r.grad_mean.cpu().numpy()
    assert np.linalg.norm(gm - want_gm) <= 2e-3 * max(np.linalg.norm(want_gm), 1e-30)


def test_shrenderer_no_grad_and_all_culled():
    cam = S.make_camera("cfg1")
    sc = S.make_scene("cfg1", seed=7, N=500)
    sc["mean"][:, 2] = -5.0  # everything behind the camera
    cfg = S.make_cfg(device=DEV, sh_order=1)
    r = S.renderer_from_scene(sc, cfg)
    with torch.no_grad():
        out = r(sc["c2w"].to(DEV), cam)
    assert out.shape == (256, 256, 3) and float(out.abs().max()) == 0.0
    assert r.total_dub_gaussians == 0


# ---------------------------------------------------------------- BASELINE-size properties (cfg 2)

def test_cfg2_full_size_properties(ops):
    """3 M Gaussians, C=4, 1297x840: no oracle at this size; check size-independent properties."""
    cam = S.make_camera("cfg2")
    sc = S.make_scene("cfg2", seed=0)
    cfg = S.make_cfg(device=DEV, sh_order=4)
    r = S.renderer_from_scene(sc, cfg)
    c2w = sc["c2w"].to(DEV)
    with torch.no_grad():
        out1 = r(c2w, cam)
        st = r._state
        ids1, s1, e1 = st["gaussian_ids"].clone(), st["start"].clone(), st["end"].clone()
        out2 = r(c2w, cam)
    n_dub = r.total_dub_gaussians
    print(f"[cfg2] n_dub = {n_dub}")
    assert torch.isfinite(out1).all() and float(out1.min()) >= 0.0 and float(out1.max()) <= 1.0 + 1e-4
    # idempotence: binning and image are run-to-run identical (deterministic emission + stable sort)
    assert torch.equal(ids1, r._state["gaussian_ids"]) and torch.equal(s1, r._state["start"])
    assert torch.equal(out1, out2)
    # ranges tile the id array exactly; keys sorted within and across tiles
    nonempty = s1 >= 0
    assert int((e1[nonempty] - s1[nonempty]).sum()) == n_dub
    order = torch.argsort(s1[nonempty])
    assert torch.equal(e1[nonempty][order][:-1], s1[nonempty][order][1:])
    depth = r.depth.view(-1)
    dbits = depth.view(torch.int32).long() & 0xFFFFFFFF
    d_sorted = dbits[ids1.long()]
    tile_of = torch.zeros(n_dub, dtype=torch.long, device=DEV)
    tile_of[s1[nonempty].long()] = 1
    seg = torch.cumsum(tile_of, 0)
    bad = (d_sorted[1:] < d_sorted[:-1]) & (seg[1:] == seg[:-1])
    assert int(bad.sum()) == 0, "depth order violated inside a tile"
    # per-Gaussian multiplicity equals its rect area
    k1_tl, k1_br = None, None
    from gaussian_splatting_3d_b200.gs.culling import tile_culling_aabb_count
    from gaussian_splatting_3d_b200.gs.renderer import project_gaussians

    mask = torch.zeros(r.N, dtype=torch.bool, device=DEV)
    normals, pts = cam.get_frustum(c2w)
    ops.culling_gaussian_bsphere(r.mean.data, r.qvec.data, r.svec.data.contiguous(), normals, pts, mask, 1.0)
    m2, cv, _, _ = project_gaussians(r.mean.data, r.qvec.data, r.svec.data.contiguous(), c2w)
    n2, tl, br = tile_culling_aabb_count(m2, cv, 16, cam, 6.0)
    area = ((br[:, 0] - tl[:, 0] + 1) * (br[:, 1] - tl[:, 1] + 1)).long() * mask.long()
    assert int(area.sum()) == n_dub
    assert torch.equal(torch.bincount(ids1.long(), minlength=r.N), area)


# ---------------------------------------------------------------- round-2 kernels against their own dense forms
def test_projection_backward_sparse_filter_equals_dense(ops):
    """K4b with the compositing backward's `touched` marks as a sparse row filter (compacting kernel, accumulate
    == 2) == the one-thread-per-Gaussian kernel over the frustum mask, on the same upstream 2-D gradients: the
    unmarked rows carry zero gradients, so both add the same values to the same rows."""
    cam = S.make_camera("cfg1")
    sc = S.make_scene("cfg1", seed=5, N=30_001)  # (not a multiple of 16 / 512: exercises the tail of the mark scan)
    N = sc["mean"].shape[0]
    p = {k: sc[k].to(DEV) for k in ("mean", "qvec", "svec_before_activation", "alpha_before_activation")}
    c2w = sc["c2w"].to(DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    marks = (torch.rand(N, device=DEV, generator=g) < 0.03).to(torch.uint8)
    frustum = torch.ones(N, dtype=torch.bool, device=DEV)
    m = marks.bool()
    gm2 = torch.zeros(N, 2, device=DEV); gcov = torch.zeros(N, 4, device=DEV); ga = torch.zeros(N, device=DEV)
    gm2[m] = torch.randn(int(m.sum()), 2, device=DEV, generator=g)
    gcov[m] = torch.randn(int(m.sum()), 4, device=DEV, generator=g)
    ga[m] = torch.randn(int(m.sum()), device=DEV, generator=g)
    outs = []
    for mask, sparse in ((frustum, False), (marks, True)):
        leaf = (torch.full((N, 3), 0.5, device=DEV), torch.full((N, 4), 0.25, device=DEV),
                torch.full((N, 3), -1.0, device=DEV), torch.full((N,), 2.0, device=DEV))
        acc = torch.zeros(N, device=DEV)
        ops.project_backward_fused(mask, p["mean"], p["qvec"], p["svec_before_activation"],
                                   p["alpha_before_activation"], 1, 1, c2w, True, gm2, gcov, ga, grad_mean_acc=acc,
                                   adc_mode=2, out=leaf, accumulate=True, sparse_filter=sparse)
        outs.append(leaf + (acc,))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert float((outs[1][0][~m] - 0.5).abs().max()) == 0.0  # unmarked rows untouched


def test_binning_specialised_passes_equal_round1_pipeline(tmp_path):
    """The specialised radix passes (binning.cu: first / last depth pass, packed rects, fused ranges) against the
    round-1 three-kernel pipeline (GS3D_SORT=classic, read once per process -> subprocesses): identical ids,
    start and end, in exact-count and in capacity mode."""
    import os
    import subprocess
    import sys

    script = r"""
import sys, torch
sys.path.insert(0, %r)
from gaussian_splatting_3d_b200 import ops, synthetic as S
dev = 'cuda:0'
cam = S.make_camera('cfg3'); sc = S.make_scene('cfg3', seed=2, N=200_003)
d = {k: v.to(dev) for k, v in sc.items() if torch.is_tensor(v)}
k1 = ops.project_cull_fused(d['mean'], d['qvec'], d['svec_before_activation'], d['alpha_before_activation'], 1, 1,
                            d['c2w'], cam, 1.0, False, 6.0, 16)
n = k1['n_dub']; nth, ntw = (cam.h + 15) // 16, (cam.w + 15) // 16
res = {}
ids = torch.empty(n, dtype=torch.int32, device=dev)
st = torch.empty(nth * ntw, dtype=torch.int32, device=dev); en = torch.empty_like(st)
keys = torch.empty(n, dtype=torch.int64, device=dev)
ops.tile_culling_aabb_start_end(k1['tl'], k1['br'], ids, st, en, k1['depth'], nth, ntw, sorted_keys=keys)
res['exact'] = [t.cpu() for t in (ids, st, en, keys)]
cap = n + 12345
ids2 = torch.full((cap,), -7, dtype=torch.int32, device=dev)
nd, ov = ops.tile_culling_aabb_start_end_capacity(k1['tl'], k1['br'], ids2, st, en, k1['depth'], nth, ntw)
res['capacity'] = [ids2[:n].cpu(), st.cpu(), en.cpu(), nd.cpu(), ov.cpu()]
small = n // 3
ids3 = torch.empty(small, dtype=torch.int32, device=dev)
nd, ov = ops.tile_culling_aabb_start_end_capacity(k1['tl'], k1['br'], ids3, st, en, k1['depth'], nth, ntw)
res['overflow'] = [ids3.cpu(), st.cpu(), en.cpu(), nd.cpu(), ov.cpu()]
torch.save(res, sys.argv[1])
""" % str(ROOT)
    outs = {}
    for mode in ("new", "classic"):
        env = dict(os.environ)
        env.pop("GS3D_SORT", None)
        if mode == "classic":
            env["GS3D_SORT"] = "classic"
        f = tmp_path / f"{mode}.pt"
        r = subprocess.run([sys.executable, "-c", script, str(f)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[mode] = torch.load(f)
    for case in ("exact", "capacity", "overflow"):
        for a, b in zip(outs["new"][case], outs["classic"][case]):
            assert torch.equal(a, b), case
    assert int(outs["new"]["overflow"][4]) == 1 and int(outs["new"]["capacity"][4]) == 0
