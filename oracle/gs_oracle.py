"""ctypes front-end of oracle/libgs_oracle.so (the C restatement in gs_oracle.c).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  numpy in, numpy out."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libgs_oracle.so"


def build(force=False):
    src = HERE / "gs_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["make", "-C", str(HERE), "-B", "libgs_oracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.gso_sigmoid.restype = C.c_float
        L.gso_sigmoid.argtypes = [C.c_float]
        L.gso_sigmoid_dsigmoid.restype = C.c_float
        L.gso_sigmoid_dsigmoid.argtypes = [C.c_float]
        L.gso_gaussian_2d.restype = C.c_float
        L.gso_gaussian_2d.argtypes = [C.c_void_p] * 3
        L.gso_gaussian_2d_backward.restype = None
        L.gso_gaussian_2d_backward.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_void_p]
        L.gso_spherical_harmonic.restype = None
        L.gso_spherical_harmonic.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.gso_calc_direction.restype = None
        L.gso_calc_direction.argtypes = [C.c_void_p] * 3
        L.gso_culling_gaussian_bsphere.restype = None
        L.gso_culling_gaussian_bsphere.argtypes = [C.c_uint32] + [C.c_void_p] * 5 + [C.c_float]
        L.gso_tile_culling_aabb_start_end.restype = C.c_int64
        L.gso_tile_culling_aabb_start_end.argtypes = [C.c_uint32] * 4 + [C.c_void_p] * 7
        L.gso_tile_based_vol_rendering_sh.restype = None
        L.gso_tile_based_vol_rendering_sh.argtypes = (
            [C.c_void_p] * 10 + [C.c_uint32] * 3 + [C.c_float] * 2 + [C.c_uint32] * 3 + [C.c_float, C.c_int]
            + [C.c_void_p] * 4)
        L.gso_tile_based_vol_rendering_backward_sh.restype = None
        L.gso_tile_based_vol_rendering_backward_sh.argtypes = (
            [C.c_uint32] + [C.c_void_p] * 15 + [C.c_uint32] * 3 + [C.c_float] * 2 + [C.c_uint32] * 3
            + [C.c_float])
        L.gso_gaussian_2d_f64.restype = C.c_float
        L.gso_gaussian_2d_f64.argtypes = [C.c_void_p] * 3
        L.gso_tile_based_vol_rendering_start_end.restype = None
        L.gso_tile_based_vol_rendering_start_end.argtypes = (
            [C.c_void_p] * 9 + [C.c_uint32] * 3 + [C.c_float] * 2 + [C.c_uint32] * 2 + [C.c_float, C.c_void_p])
        L.gso_tile_based_vol_rendering_backward_start_end.restype = None
        L.gso_tile_based_vol_rendering_backward_start_end.argtypes = (
            [C.c_uint32] + [C.c_void_p] * 14 + [C.c_uint32] * 3 + [C.c_float] * 2 + [C.c_uint32] * 2 + [C.c_float])
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sigmoid(x):
    return float(lib().gso_sigmoid(x))


def sigmoid_dsigmoid(s):
    return float(lib().gso_sigmoid_dsigmoid(s))


def gaussian_2d(mean, cov, query):
    m, c, q = _f32(mean), _f32(cov).reshape(-1), _f32(query)
    return float(lib().gso_gaussian_2d(_p(m), _p(c), _p(q)))


def gaussian_2d_backward(mean, cov, query, grad):
    m, c, q = _f32(mean), _f32(cov).reshape(-1), _f32(query)
    out = np.zeros(6, dtype=np.float64)
    lib().gso_gaussian_2d_backward(_p(m), _p(c), _p(q), float(grad), _p(out))
    return out[:2].copy(), out[2:].reshape(2, 2).copy()


def spherical_harmonic(direction, C_):
    d = _f32(direction)
    out = np.zeros(16, dtype=np.float32)
    lib().gso_spherical_harmonic(_p(d), _p(out), int(C_))
    return out[: C_ * C_].copy()


def calc_direction(pos3, c2w):
    p, c = _f32(pos3), _f32(c2w).reshape(-1)
    out = np.zeros(3, dtype=np.float32)
    lib().gso_calc_direction(_p(out), _p(p), _p(c))
    return out


def culling_gaussian_bsphere(mean, svec, normal, pts, thresh):
    mean, svec, normal, pts = _f32(mean), _f32(svec), _f32(normal), _f32(pts)
    mask = np.zeros(mean.shape[0], dtype=np.uint8)
    lib().gso_culling_gaussian_bsphere(mean.shape[0], _p(mean), _p(svec), _p(normal), _p(pts), _p(mask),
                                       float(thresh))
    return mask.astype(bool)


def tile_culling_aabb_start_end(aabb_tl, aabb_br, depth, n_dub, n_tiles_h, n_tiles_w):
    """-> (gaussian_ids int32 [n_dub], start, end int32 [n_tiles], sorted_keys int64 [n_dub])."""
    tl, br, depth = _i32(aabb_tl), _i32(aabb_br), _f32(depth).reshape(-1)
    ids = np.zeros(max(n_dub, 1), dtype=np.int32)
    keys = np.zeros(max(n_dub, 1), dtype=np.int64)
    start = np.full(n_tiles_h * n_tiles_w, -1, dtype=np.int32)
    end = np.full(n_tiles_h * n_tiles_w, -1, dtype=np.int32)
    r = lib().gso_tile_culling_aabb_start_end(tl.shape[0], int(n_dub), int(n_tiles_h), int(n_tiles_w),
                                              _p(tl), _p(br), _p(depth), _p(ids), _p(start), _p(end),
                                              _p(keys))
    if r < 0:
        raise RuntimeError("oracle: emitted duplicate count != n_dub (aabb_culling.h:228)")
    return ids[:n_dub], start, end, keys[:n_dub]


def render_sh_forward(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, tile_size,
                      n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C_, thresh, bg_rgb=None,
                      diagnostics=False):
    mean, cov, sh, alpha = _f32(mean), _f32(cov).reshape(-1, 4), _f32(sh_coeffs), _f32(alpha)
    start, end, ids = _i32(start), _i32(end), _i32(gaussian_ids)
    if ids.size == 0:
        ids = np.zeros(1, dtype=np.int32)
    topleft, c2w = _f32(topleft), _f32(c2w).reshape(-1)
    out = np.zeros(H * W * 3, dtype=np.float32)
    bg = None if bg_rgb is None else _f32(bg_rgb)
    fT = np.ones(H * W, dtype=np.float32) if diagnostics else None
    nc = np.zeros(H * W, dtype=np.int32) if diagnostics else None
    mg = np.ones(H * W, dtype=np.float32) if diagnostics else None
    lib().gso_tile_based_vol_rendering_sh(
        _p(mean), _p(cov), _p(sh), _p(alpha), _p(start), _p(end), _p(ids), _p(out), _p(topleft), _p(c2w),
        int(tile_size), int(n_tiles_h), int(n_tiles_w), float(pixel_size_x), float(pixel_size_y), int(H),
        int(W), int(C_), float(thresh), 0 if bg is None else 1, _p(bg), _p(fT), _p(nc), _p(mg))
    if diagnostics:
        return out, fT, nc, mg
    return out


def render_sh_backward(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_out, topleft, c2w,
                       tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C_, thresh):
    mean, cov, sh, alpha = _f32(mean), _f32(cov).reshape(-1, 4), _f32(sh_coeffs), _f32(alpha)
    start, end, ids = _i32(start), _i32(end), _i32(gaussian_ids)
    if ids.size == 0:
        ids = np.zeros(1, dtype=np.int32)
    out, grad_out = _f32(out).reshape(-1), _f32(grad_out).reshape(-1)
    topleft, c2w = _f32(topleft), _f32(c2w).reshape(-1)
    N = mean.shape[0]
    g_mean = np.zeros((N, 2), dtype=np.float32)
    g_cov = np.zeros((N, 4), dtype=np.float32)
    g_sh = np.zeros((N, 3, C_ * C_), dtype=np.float32)
    g_alpha = np.zeros(N, dtype=np.float32)
    lib().gso_tile_based_vol_rendering_backward_sh(
        N, _p(mean), _p(cov), _p(sh), _p(alpha), _p(start), _p(end), _p(ids), _p(out), _p(g_mean),
        _p(g_cov), _p(g_sh), _p(g_alpha), _p(grad_out), _p(topleft), _p(c2w), int(tile_size),
        int(n_tiles_h), int(n_tiles_w), float(pixel_size_x), float(pixel_size_y), int(H), int(W), int(C_),
        float(thresh))
    return g_mean, g_cov, g_sh, g_alpha


# ---------------------------------------------------------------- legacy RGB path (8f rank 2)

def gaussian_2d_f64(mean, cov, query):
    mean, cov, query = _f32(mean), _f32(cov).reshape(-1), _f32(query)
    return float(lib().gso_gaussian_2d_f64(_p(mean), _p(cov), _p(query)))


def render_rgb_forward(mean, cov, color, alpha, start, end, gaussian_ids, topleft, tile_size, n_tiles_h,
                       n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh, diagnostics=False):
    """vol_render.h:716-798: RGB compositing over (start, end) ranges; -> out [H*W*3] (+ margin)."""
    mean, cov, color, alpha = _f32(mean), _f32(cov).reshape(-1, 4), _f32(color), _f32(alpha)
    start, end, ids = _i32(start), _i32(end), _i32(gaussian_ids)
    if ids.size == 0:
        ids = np.zeros(1, dtype=np.int32)
    topleft = _f32(topleft)
    out = np.zeros(H * W * 3, dtype=np.float32)
    mg = np.ones(H * W, dtype=np.float32) if diagnostics else None
    lib().gso_tile_based_vol_rendering_start_end(
        _p(mean), _p(cov), _p(color), _p(alpha), _p(start), _p(end), _p(ids), _p(out), _p(topleft),
        int(tile_size), int(n_tiles_h), int(n_tiles_w), float(pixel_size_x), float(pixel_size_y), int(H), int(W),
        float(thresh), _p(mg))
    return (out, mg) if diagnostics else out


def render_rgb_backward(mean, cov, color, alpha, start, end, gaussian_ids, out, grad_out, topleft, tile_size,
                        n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh):
    """vol_render.h:800-923 -> (grad_mean [N,2], grad_cov [N,4], grad_color [N,3], grad_alpha [N])."""
    mean, cov, color, alpha = _f32(mean), _f32(cov).reshape(-1, 4), _f32(color), _f32(alpha)
    start, end, ids = _i32(start), _i32(end), _i32(gaussian_ids)
    if ids.size == 0:
        ids = np.zeros(1, dtype=np.int32)
    out, grad_out, topleft = _f32(out).reshape(-1), _f32(grad_out).reshape(-1), _f32(topleft)
    N = mean.shape[0]
    g_mean = np.zeros((N, 2), dtype=np.float32)
    g_cov = np.zeros((N, 4), dtype=np.float32)
    g_color = np.zeros((N, 3), dtype=np.float32)
    g_alpha = np.zeros(N, dtype=np.float32)
    lib().gso_tile_based_vol_rendering_backward_start_end(
        N, _p(mean), _p(cov), _p(color), _p(alpha), _p(start), _p(end), _p(ids), _p(out), _p(g_mean), _p(g_cov),
        _p(g_color), _p(g_alpha), _p(grad_out), _p(topleft), int(tile_size), int(n_tiles_h), int(n_tiles_w),
        float(pixel_size_x), float(pixel_size_y), int(H), int(W), float(thresh))
    return g_mean, g_cov, g_color, g_alpha
