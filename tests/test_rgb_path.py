"""Legacy RGB path: tile_based_vol_rendering_start_end (+ backward), render_start_end
and GaussianRenderer.

CPU: the oracle's RGB restatement is pinned by the reference's KAT for kernel_gaussian_2d, by finite
differences, and by the (reference-pinned) SH restatement, to which it must reduce for degree-0 colours.
GPU (-m gpu): the RGB-mode kernels through the C ABI against the oracle, against the REAL reference
extension (oracle/_ref), and the module against the torch-level reference flow.
Tolerances: images 1e-4 max-abs, gradients 1e-3 relative (like-for-like against the extension).
"""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from gaussian_splatting_3d_b200 import synthetic as S

ROOT = Path(__file__).resolve().parent.parent
DEV = "cuda:0"
Y00 = 0.28209479177387814


def _aux(name, seed, n):
    from oracle import ref_torch as R

    cam = S.make_camera(name)
    sc = S.make_scene(name, seed=seed, N=n, C=1)
    p = {k: sc[k] for k in ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")}
    img, aux = R.reference_forward(p, sc["c2w"], cam, 1, T_thresh=1e-4, return_aux=True)
    H, W = cam.h, cam.w
    consts = (16, (H + 15) // 16, (W + 15) // 16, np.float32(1 / cam.fx), np.float32(1 / cam.fy), H, W, 1e-4)
    topleft = np.array([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=np.float32)
    d = dict(mean=aux["mean2d"].detach().numpy().astype(np.float32),
             cov=aux["cov"].detach().numpy().reshape(-1, 4).astype(np.float32),
             alpha=aux["alpha"].detach().numpy().astype(np.float32),
             color=torch.sigmoid(aux["sh"][..., 0].detach() * Y00).numpy().astype(np.float32),
             start=aux["start"], end=aux["end"], ids=aux["ids"], topleft=topleft, consts=consts)
    return sc, cam, img, d


# ------------------------------------------------------------------------------------------ CPU

def test_kat_gaussian_f64(golden_dir):
    import json

    from oracle import gs_oracle as K

    kat = json.loads((golden_dir / "kat.json").read_text())
    # kernel_gaussian_2d == kernel_gaussian_2d_float == 0.951229393 on test/gaussian_test.py's vector
    v = K.gaussian_2d_f64([0.1, 0.2], [0.5, 0.2, 0.2, 0.8], [0.3, 0.4])
    assert abs(v - kat["kernel_gaussian_2d_float"]) < 1e-7
    assert abs(v - kat["gaussian_test"]["G"]) < 1e-7
    assert K.gaussian_2d_f64([0, 0], [1, 2, 2, 1], [1, 0]) == pytest.approx(np.exp(-500.0), abs=1e-30)  # radial < 0


def test_oracle_rgb_reduces_to_the_sh_restatement_for_degree_zero_colours():
    from oracle import gs_oracle as K

    sc, cam, img, d = _aux("cfg1", 4, 3000)
    out, margin = K.render_rgb_forward(d["mean"], d["cov"], d["color"], d["alpha"], d["start"], d["end"], d["ids"],
                                       d["topleft"], *d["consts"], diagnostics=True)
    err = np.abs(out.reshape(-1, 3) - img.detach().numpy().reshape(-1, 3)).max(axis=1)
    stable = margin > 2e-3  # FP64 vs FP32 Gaussian: decisions within 0.2 % of 1/255 may differ
    assert stable.mean() > 0.98 and err[stable].max() <= 1e-5


def test_oracle_rgb_backward_matches_finite_differences():
    from oracle import gs_oracle as K

    sc, cam, img, d = _aux("cfg1", 6, 400)
    rng = np.random.default_rng(1)
    H, W = cam.h, cam.w
    g_out = rng.standard_normal(H * W * 3).astype(np.float32)
    args = (d["start"], d["end"], d["ids"])

    def f(m, c, col, a):
        o = K.render_rgb_forward(m, c, col, a, *args, d["topleft"], *d["consts"])
        return float(o.astype(np.float64) @ g_out.astype(np.float64))

    out = K.render_rgb_forward(d["mean"], d["cov"], d["color"], d["alpha"], *args, d["topleft"], *d["consts"])
    gm, gc, gcol, ga = K.render_rgb_backward(d["mean"], d["cov"], d["color"], d["alpha"], *args, out, g_out,
                                             d["topleft"], *d["consts"])
    touched = np.flatnonzero(np.abs(gcol).sum(axis=1) > 0)
    assert touched.size > 20
    ok = n = 0
    for which, arr, grad, eps in (("mean", d["mean"], gm, 2e-4), ("alpha", d["alpha"], ga, 1e-3),
                                  ("color", d["color"], gcol, 1e-2)):
        per = arr.reshape(arr.shape[0], -1).shape[1]
        for g in rng.choice(touched, size=12, replace=False):
            i = g * per + int(rng.integers(per))
            flat = arr.reshape(-1)
            if which == "alpha" and flat[i] > 0.98:
                continue
            hi, lo = flat.copy(), flat.copy()
            hi[i] += eps
            lo[i] -= eps
            pk = lambda v: {"mean": (v.reshape(arr.shape), d["cov"], d["color"], d["alpha"]),  # noqa: E731
                            "alpha": (d["mean"], d["cov"], d["color"], v),
                            "color": (d["mean"], d["cov"], v.reshape(arr.shape), d["alpha"])}[which]
            fd = (f(*pk(hi)) - f(*pk(lo))) / (2 * eps)
            gi = grad.reshape(-1)[i]
            n += 1
            ok += abs(fd - gi) <= 3e-2 * max(abs(fd), abs(gi)) + 2e-3 * np.abs(grad).max()
    assert ok >= 0.85 * n, (ok, n)  # a perturbation can flip a 1/255 decision somewhere: not every probe is smooth


def test_rgb_bindings_reject_cpu_tensors():
    import gaussian_splatting_3d_b200._gs as gs

    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA"):
        gs.tile_based_vol_rendering_start_end(z(2, 2), z(2, 4), z(2, 3), z(2), z(1, dtype=torch.int32),
                                              z(1, dtype=torch.int32), z(2, dtype=torch.int32), z(16 * 16 * 3), z(2),
                                              16, 1, 1, 0.1, 0.1, 16, 16, 1e-4)


# ------------------------------------------------------------------------------------------ GPU

def _dev(d):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    return {k: (t(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,n", [("cfg1", 0, 10_000), ("cfg3", 2, 30_000)])
def test_gpu_rgb_forward_backward_vs_oracle(name, seed, n):
    import gaussian_splatting_3d_b200._gs as gs
    from oracle import gs_oracle as K

    sc, cam, img, d = _aux(name, seed, n)
    H, W = cam.h, cam.w
    args = (d["start"], d["end"], d["ids"])
    want, margin = K.render_rgb_forward(d["mean"], d["cov"], d["color"], d["alpha"], *args, d["topleft"], *d["consts"],
                                        diagnostics=True)
    g = _dev(d)
    consts = tuple(float(c) if isinstance(c, np.floating) else c for c in d["consts"])
    out = torch.zeros(H * W * 3, device=DEV)
    gs.tile_based_vol_rendering_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"], g["end"], g["ids"],
                                          out, g["topleft"], *consts)
    err = np.abs(out.cpu().numpy().reshape(-1, 3) - want.reshape(-1, 3)).max(axis=1)
    stable = margin > 2e-3
    assert stable.mean() > 0.98
    assert err[stable].max() <= 1e-4, err[stable].max()          # images: 1e-4 max-abs
    assert err.max() <= 2.5 / 255                                 # fragile pixels: at most two flipped splats
    tgt = S.make_target(cam, seed).numpy().reshape(-1)
    g_out = (2.0 * (want - tgt) / want.size).astype(np.float32)
    wg = K.render_rgb_backward(d["mean"], d["cov"], d["color"], d["alpha"], *args, want, g_out, d["topleft"],
                               *d["consts"])
    gm, gc, gcol, ga = (torch.zeros_like(g["mean"]), torch.zeros_like(g["cov"]), torch.zeros_like(g["color"]),
                        torch.zeros_like(g["alpha"]))
    gs.tile_based_vol_rendering_backward_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"], g["end"],
                                                   g["ids"], out, gm, gc, gcol, ga, torch.from_numpy(g_out).to(DEV),
                                                   g["topleft"], *consts)
    for tag, got, w in (("mean", gm, wg[0]), ("cov", gc, wg[1]), ("color", gcol, wg[2]), ("alpha", ga, wg[3])):
        a, b = got.cpu().numpy().astype(np.float64).ravel(), w.astype(np.float64).ravel()
        l2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        # vs the CPU oracle a handful of skip decisions differ (glibc vs CUDA exp): 3e-3 in L2 here, the 1e-3
        # bound is enforced like for like against the real extension below
        assert l2 <= 3e-3, (tag, l2)


def _load_ref():
    sos = sorted((ROOT / "oracle" / "_ref").glob("_gs_ref*.so"))
    if not sos:
        return None
    spec = importlib.util.spec_from_file_location("_gs_ref", sos[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,n", [("cfg1", 0, 10_000), ("cfg2", 2, 200_000)])
def test_gpu_rgb_matches_reference_extension(name, seed, n):
    import gaussian_splatting_3d_b200._gs as ours

    try:
        ref = _load_ref()
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference extension not loadable: {e}")
    if ref is None:
        pytest.skip("oracle/_ref/_gs_ref*.so not present")
    sc, cam, img, d = _aux(name, seed, n)
    H, W = cam.h, cam.w
    g = _dev(d)
    consts = tuple(float(c) if isinstance(c, np.floating) else c for c in d["consts"])

    def fwd(mod):
        out = torch.zeros(H * W * 3, device=DEV)
        mod.tile_based_vol_rendering_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"], g["end"],
                                               g["ids"], out, g["topleft"], *consts)
        torch.cuda.synchronize()
        return out

    out_r, out_o = fwd(ref), fwd(ours)
    err = (out_r - out_o).abs()
    assert float(err.max()) <= 1e-4, f"image differs from the reference extension: {float(err.max()):.3e}"
    tgt = S.make_target(cam, seed).to(DEV).reshape(-1)
    g_out = (2.0 * (out_r - tgt) / out_r.numel()).contiguous()

    def bwd(mod, out):
        gm, gc, gcol, ga = (torch.zeros_like(g["mean"]), torch.zeros_like(g["cov"]), torch.zeros_like(g["color"]),
                            torch.zeros_like(g["alpha"]))
        mod.tile_based_vol_rendering_backward_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"],
                                                        g["end"], g["ids"], out, gm, gc, gcol, ga, g_out, g["topleft"],
                                                        *consts)
        torch.cuda.synchronize()
        return gm, gc, gcol, ga

    for tag, a, b in zip(("mean", "cov", "color", "alpha"), bwd(ours, out_o), bwd(ref, out_r)):
        a64, b64 = a.double().reshape(-1), b.double().reshape(-1)
        l2 = float((a64 - b64).norm() / b64.norm().clamp_min(1e-30))
        mx = float((a64 - b64).abs().max() / b64.abs().max().clamp_min(1e-30))
        print(f"[ref-ext rgb {name}] grad_{tag}: L2 rel {l2:.2e}, max rel {mx:.2e}")
        assert l2 <= 1e-3 and mx <= 1e-3, f"grad_{tag}: L2 {l2:.3e} max {mx:.3e}"  # gradients: 1e-3 relative


@pytest.mark.gpu
def test_gpu_gaussian_renderer_module_matches_reference_flow():
    """GaussianRenderer.forward + backward (renderer.py:1219-1305) vs the torch-level reference flow with the
    oracle's RGB kernels; then one adaptive-control round trip (split, prune) keeps rendering."""
    from gaussian_splatting_3d_b200.gs.renderer import GaussianRenderer, render_start_end
    from oracle import gs_oracle as K

    sc, cam, img, d = _aux("cfg1", 3, 8000)
    cfg = S.make_cfg(device=DEV, sh_order=1, color_act="sigmoid", alpha_init=0.5, svec_init=0.01,
                     adaptive_control_iteration=1, alpha_reset_period=0, pos_grad_thresh=1e-7, alpha_thresh=0.05)
    r = GaussianRenderer(cfg, sc["mean"].clone(), torch.full((sc["mean"].shape[0], 3), 0.5))
    r._set([sc["mean"].to(DEV), sc["qvec"].to(DEV), sc["svec_before_activation"].to(DEV),
            (sc["sh_coeffs"][..., 0] * Y00).to(DEV).contiguous(), sc["alpha_before_activation"].to(DEV)])
    out = r(sc["c2w"].to(DEV), cam)
    assert r.total_dub_gaussians == int(d["ids"].shape[0])
    H, W = cam.h, cam.w
    want, margin = K.render_rgb_forward(d["mean"], d["cov"], d["color"], d["alpha"], d["start"], d["end"], d["ids"],
                                        d["topleft"], *d["consts"], diagnostics=True)
    err = np.abs(out.detach().cpu().numpy().reshape(-1, 3) - want.reshape(-1, 3)).max(axis=1)
    stable = margin > 2e-3
    assert err[stable].max() <= 1e-4 and err.max() <= 2.5 / 255
    tgt = S.make_target(cam, 3).to(DEV)
    ((out - tgt) ** 2).mean().backward()
    for n_ in GaussianRenderer._NAMES:
        gr = getattr(r, n_).grad
        assert gr is not None and torch.isfinite(gr).all() and float(gr.abs().max()) > 0, n_
    # the standalone Function gives the same image from the same projected inputs
    g = _dev(d)
    consts = tuple(float(c) if isinstance(c, np.floating) else c for c in d["consts"])
    out2 = render_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"], g["end"], g["ids"], g["topleft"],
                            *consts)
    assert float((out2 - out.detach().reshape(-1)).abs().max()) <= 1e-4
    n0 = r.N
    r.split_gaussians()
    assert r.N > n0
    r.remove_low_alpha_gaussians()
    out3 = r(sc["c2w"].to(DEV), cam)
    assert out3.shape == (H, W, 3) and torch.isfinite(out3).all()


_CSR_SCRIPT = r"""
import importlib.util, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(%r)
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
from gaussian_splatting_3d_b200 import synthetic as S, ops
import gaussian_splatting_3d_b200._gs as ours
import test_rgb_path as T
DEV = 'cuda:0'
ref = T._load_ref()
sc, cam, img, d = T._aux('cfg1', 6, 10_000)
H, W = cam.h, cam.w
g = T._dev(d)
consts = tuple(float(c) if isinstance(c, np.floating) else c for c in d['consts'])
counts = (g['end'] - g['start']).clamp(min=0)
offset = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=DEV)
offset[1:] = torch.cumsum(counts, 0)
res, outs = {}, {}
for name in ('tile_based_vol_rendering', 'tile_based_vol_rendering_v1', 'tile_based_vol_rendering_v2'):
    for tag, mod in (('ref', ref), ('ours', ours)):
        out = torch.zeros(H * W * 3, device=DEV)
        getattr(mod, name)(g['mean'], g['cov'], g['color'], g['alpha'], offset, g['ids'], out, g['topleft'], *consts)
        torch.cuda.synchronize()
        outs[(name, tag)] = out
    res[name] = float((outs[(name, 'ref')] - outs[(name, 'ours')]).abs().max())
out_r = outs[('tile_based_vol_rendering', 'ref')]
tgt = S.make_target(cam, 6).to(DEV).reshape(-1)
g_out = (2.0 * (out_r - tgt) / out_r.numel()).contiguous()
grads = {}
for tag, mod in (('ref', ref), ('ours', ours)):
    gm, gc, gcol, ga = (torch.zeros_like(g['mean']), torch.zeros_like(g['cov']), torch.zeros_like(g['color']),
                        torch.zeros_like(g['alpha']))
    mod.tile_based_vol_rendering_backward(g['mean'], g['cov'], g['color'], g['alpha'], offset, g['ids'],
                                          outs[('tile_based_vol_rendering', tag)], gm, gc, gcol, ga, g_out,
                                          g['topleft'], *consts)
    torch.cuda.synchronize()
    grads[tag] = (gm, gc, gcol, ga)
for tag, a, b in zip(('mean', 'cov', 'color', 'alpha'), grads['ours'], grads['ref']):
    a64, b64 = a.double().reshape(-1), b.double().reshape(-1)
    res['grad_' + tag] = float((a64 - b64).norm() / b64.norm().clamp_min(1e-30))
print('CSR_RESULT ' + json.dumps(res))
"""


@pytest.mark.gpu
def test_gpu_csr_offset_bindings_match_reference_extension():
    """The deprecated CSR-`offset` bindings (bindings.cpp:15-25: tile_based_vol_rendering{,_v1,_v2} and the
    backward; tile_culling_aabb) as adapters over the start/end kernels.  The rendering variants are compared
    with the REAL reference extension on a CSR built from the same tile lists -- in a subprocess: the reference's
    deprecated kernels are not exercised by any reference caller any more, and a fault inside them must not poison
    this process's CUDA context."""
    import json
    import subprocess
    import sys

    import gaussian_splatting_3d_b200._gs as ours

    # (1) the adapters against this repo's start/end bindings on the same lists: the same kernel, so bit-identical
    sc, cam, img, d = _aux("cfg1", 6, 10_000)
    H, W = cam.h, cam.w
    g = _dev(d)
    consts = tuple(float(c) if isinstance(c, np.floating) else c for c in d["consts"])
    counts = (g["end"] - g["start"]).clamp(min=0)
    offset = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=DEV)
    offset[1:] = torch.cumsum(counts, 0)
    want = torch.zeros(H * W * 3, device=DEV)
    ours.tile_based_vol_rendering_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"], g["end"], g["ids"],
                                            want, g["topleft"], *consts)
    for name in ("tile_based_vol_rendering", "tile_based_vol_rendering_v1", "tile_based_vol_rendering_v2"):
        got = torch.zeros(H * W * 3, device=DEV)
        getattr(ours, name)(g["mean"], g["cov"], g["color"], g["alpha"], offset, g["ids"], got, g["topleft"], *consts)
        assert torch.equal(got, want), name
    g_out = torch.rand(H * W * 3, device=DEV) * 1e-6
    both = []
    for csr in (False, True):
        gm, gc, gcol, ga = (torch.zeros_like(g["mean"]), torch.zeros_like(g["cov"]), torch.zeros_like(g["color"]),
                            torch.zeros_like(g["alpha"]))
        if csr:
            ours.tile_based_vol_rendering_backward(g["mean"], g["cov"], g["color"], g["alpha"], offset, g["ids"], want,
                                                   gm, gc, gcol, ga, g_out, g["topleft"], *consts)
        else:
            ours.tile_based_vol_rendering_backward_start_end(g["mean"], g["cov"], g["color"], g["alpha"], g["start"],
                                                             g["end"], g["ids"], want, gm, gc, gcol, ga, g_out,
                                                             g["topleft"], *consts)
        both.append((gm, gc, gcol, ga))
    for a, b in zip(*both):  # (float atomics: order differs from launch to launch)
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12
    # (2) against the REAL reference extension, when its deprecated kernels run at all on this GPU (on the B200 box
    # they fault with an illegal address -- no reference caller exercises them any more): subprocess
    if _load_ref() is not None:
        r = subprocess.run([sys.executable, "-c", _CSR_SCRIPT % str(ROOT)], capture_output=True, text=True,
                           timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("CSR_RESULT ")]
        if r.returncode == 0 and line:
            res = json.loads(line[0][len("CSR_RESULT "):])
            print("[csr vs reference extension]", res)
            for k, v in res.items():
                assert v <= (1e-3 if k.startswith("grad_") else 1e-4), (k, v)
        else:
            print("[csr] the reference's deprecated CSR kernels did not run here:", r.stderr[-300:].replace("\n", " "))
    # (3) tile_culling_aabb: proper CSR of the same binning as the start/end form
    from gaussian_splatting_3d_b200 import ops

    p = {k: sc[k].to(DEV) for k in ("mean", "qvec", "svec_before_activation", "alpha_before_activation")}
    k1 = ops.project_cull_fused(p["mean"], p["qvec"], p["svec_before_activation"], p["alpha_before_activation"], 1, 1,
                                sc["c2w"].to(DEV), cam, 1.0, False, 6.0, 16)
    nth, ntw = (H + 15) // 16, (W + 15) // 16
    ids = torch.empty(k1["n_dub"], dtype=torch.int32, device=DEV)
    off = torch.empty(nth * ntw + 1, dtype=torch.int32, device=DEV)
    ours.tile_culling_aabb(k1["tl"], k1["br"], ids, off, k1["depth"], nth, ntw)
    ids2 = torch.empty_like(ids)
    st = torch.empty(nth * ntw, dtype=torch.int32, device=DEV)
    en = torch.empty_like(st)
    ours.tile_culling_aabb_start_end(k1["tl"], k1["br"], ids2, st, en, k1["depth"], nth, ntw)
    assert torch.equal(ids, ids2) and int(off[-1]) == k1["n_dub"] and int(off[0]) == 0
    nz = st >= 0
    assert torch.equal(off[:-1][nz], st[nz]) and torch.equal(off[1:][nz], en[nz])
    assert bool((off[1:] >= off[:-1]).all())
