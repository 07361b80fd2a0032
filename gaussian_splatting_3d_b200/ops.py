"""Tensor-level wrappers over the C ABI (capi).  Validation mirrors the reference's CHECK_* macros
(gs/src/include/common.h:29-54): a wrong device / layout / dtype raises RuntimeError.  Scratch
memory comes from torch's caching allocator (the reference cudaMalloc/cudaFree'd inside the call,
aabb_culling.h:204-259).  Everything is enqueued on the current torch stream.
"""
import ctypes as C

import torch

from . import capi
from .capi import check, ptr

_F32, _I32, _BOOL, _U8, _I64 = torch.float32, torch.int32, torch.bool, torch.uint8, torch.int64
_I64P_t = C.POINTER(C.c_int64)


def _chk(t, name, dtype):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    dts = dtype if isinstance(dtype, tuple) else (dtype,)
    if t.dtype not in dts:
        kind = {_F32: "a floating", _I32: "an int", _BOOL: "an bool"}.get(dts[0], str(dts[0]))
        raise RuntimeError(f"{name} must be {kind} tensor")
    return t


def _stream(t):
    return capi.current_stream(t.device)


_SCRATCH = {}  # (device, stream) -> persistent uint8 buffer, grown on demand


def _scratch(nbytes, device):
    """Kernel scratch.  Large requests (the binning's ping-pong buffers, ~200 MB at cfg 2) are served from one
    persistent buffer per (device, stream) -- consecutive calls on a stream are ordered, so they can share it --
    instead of a fresh caching-allocator block per call.  During CUDA-graph capture the buffer comes from the
    graph's own pool (it must stay alive with the graph, not with this cache)."""
    nbytes = max(int(nbytes), 256)
    if nbytes < (1 << 20) or torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=_U8, device=device)
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        _SCRATCH[key] = buf = torch.empty(nbytes + nbytes // 8, dtype=_U8, device=device)
    return buf


# ---------------------------------------------------------------- a1
def get_frustum(c2w, camera_info):
    _chk(c2w, "c2w", _F32)
    normals = torch.empty(6, 3, dtype=_F32, device=c2w.device)
    pts = torch.empty(6, 3, dtype=_F32, device=c2w.device)
    cam = capi.camera_struct(camera_info)
    check(capi.lib.gs3d_get_frustum(ptr(c2w), C.byref(cam), ptr(normals), ptr(pts), _stream(c2w)),
          "get_frustum")
    return normals, pts


# ---------------------------------------------------------------- a2
def culling_gaussian_bsphere(mean, qvec, svec, normal, pts, mask, thresh):
    for t, n in ((mean, "mean"), (qvec, "qvec"), (svec, "svec"), (normal, "normal"), (pts, "pts")):
        _chk(t, n, _F32)
    _chk(mask, "mask", _BOOL)
    check(capi.lib.gs3d_culling_gaussian_bsphere(mean.size(0), ptr(mean), ptr(qvec), ptr(svec),
                                                 ptr(normal), ptr(pts), ptr(mask), float(thresh),
                                                 _stream(mean)), "culling_gaussian_bsphere")


# ---------------------------------------------------------------- a4
def project_gaussians_forward(mean, qvec, svec, c2w, want_JW=True):
    for t, n in ((mean, "mean"), (qvec, "qvec"), (svec, "svec"), (c2w, "c2w")):
        _chk(t, n, _F32)
    N = mean.size(0)
    dev = mean.device
    mean2d = torch.empty(N, 2, dtype=_F32, device=dev)
    cov = torch.empty(N, 2, 2, dtype=_F32, device=dev)
    JW = torch.empty(N, 3, 3, dtype=_F32, device=dev) if want_JW else None
    depth = torch.empty(N, 1, dtype=_F32, device=dev)
    check(capi.lib.gs3d_project_gaussians(N, ptr(mean), ptr(qvec), ptr(svec), ptr(c2w), ptr(mean2d),
                                          ptr(cov), ptr(JW), ptr(depth), _stream(mean)),
          "project_gaussians")
    return mean2d, cov, JW, depth


def project_gaussians_backward(mean, qvec, svec, c2w, g_mean2d, g_cov, g_depth, detach_depth):
    N = mean.size(0)
    dev = mean.device
    gm = torch.empty(N, 3, dtype=_F32, device=dev)
    gq = torch.empty(N, 4, dtype=_F32, device=dev)
    gs = torch.empty(N, 3, dtype=_F32, device=dev)
    g_mean2d = _chk(g_mean2d.contiguous(), "grad_mean2d", _F32)
    g_cov = _chk(g_cov.contiguous(), "grad_cov", _F32)
    if g_depth is not None:
        g_depth = _chk(g_depth.contiguous(), "grad_depth", _F32)
    check(capi.lib.gs3d_project_gaussians_backward(N, ptr(mean), ptr(qvec), ptr(svec), ptr(c2w),
                                                   ptr(g_mean2d), ptr(g_cov), ptr(g_depth),
                                                   1 if detach_depth else 0, ptr(gm), ptr(gq),
                                                   ptr(gs), _stream(mean)),
          "project_gaussians_backward")
    return gm, gq, gs


# ---------------------------------------------------------------- a5
def tile_culling_aabb_count(mean2d, cov, tile_size, camera_info, D):
    _chk(mean2d, "mean", _F32)
    _chk(cov, "cov", _F32)
    N = mean2d.size(0)
    dev = mean2d.device
    tl = torch.empty(N, 2, dtype=_I32, device=dev)
    br = torch.empty(N, 2, dtype=_I32, device=dev)
    n = C.c_int64(0)
    scratch = _scratch(capi.lib.gs3d_count_scratch_bytes(N), dev)
    cam = capi.camera_struct(camera_info)
    check(capi.lib.gs3d_tile_culling_aabb_count(N, ptr(mean2d), ptr(cov), int(tile_size),
                                                C.byref(cam), float(D), ptr(tl), ptr(br),
                                                C.byref(n), ptr(scratch), scratch.numel(),
                                                _stream(mean2d)), "tile_culling_aabb_count")
    return int(n.value), tl, br


# ---------------------------------------------------------------- fused K1
def project_cull_fused(mean, qvec, svec_param, alpha_param, svec_act, alpha_act, c2w, camera_info,
                       frustum_radius, skip_frustum_culling, tile_D, tile_size, cnt=None,
                       want_records=True, want_activated=True, sync_count=True, want_projection=True):
    """sync_count=False: no host read-back of the duplicate count (out["n_dub"] is None, out["n_dub_dev"] an
    int64 [1] device view of it): the call, and the step around it, can be captured in a CUDA graph.
    want_projection=False (needs want_records): mean2d / cov are not written as separate arrays; out["mean2d"]
    and out["cov"] are then strided VIEWS into the staging records (floats 0-1 and 8-11)."""
    for t, n in ((mean, "mean"), (qvec, "qvec"), (svec_param, "svec"), (alpha_param, "alpha"),
                 (c2w, "c2w")):
        _chk(t, n, _F32)
    if cnt is not None:
        _chk(cnt, "cnt", _I32)
    if not want_projection and not want_records:
        raise RuntimeError("want_projection=False needs want_records=True")
    N = mean.size(0)
    dev = mean.device
    out = {
        "mask": torch.empty(N, dtype=_BOOL, device=dev),
        "mean2d": torch.empty(N, 2, dtype=_F32, device=dev) if want_projection else None,
        "cov": torch.empty(N, 2, 2, dtype=_F32, device=dev) if want_projection else None,
        "depth": torch.empty(N, 1, dtype=_F32, device=dev),
        "tl": torch.empty(N, 2, dtype=_I32, device=dev),
        "br": torch.empty(N, 2, dtype=_I32, device=dev),
        "records": torch.empty(N, 12, dtype=_F32, device=dev) if want_records else None,
        "svec": torch.empty(N, 3, dtype=_F32, device=dev) if want_activated else None,
        "alpha": torch.empty(N, dtype=_F32, device=dev) if want_activated else None,
    }
    n = C.c_int64(0)
    scratch = _scratch(256, dev)
    cam = capi.camera_struct(camera_info)
    n_arg = C.cast(C.pointer(n), C.c_void_p) if sync_count else None
    check(capi.lib.gs3d_project_cull_fused(
        N, ptr(mean), ptr(qvec), ptr(svec_param), ptr(alpha_param), int(svec_act), int(alpha_act),
        ptr(c2w), C.byref(cam), float(frustum_radius), 1 if skip_frustum_culling else 0,
        float(tile_D), int(tile_size), ptr(out["mask"]), ptr(out["mean2d"]), ptr(out["cov"]),
        ptr(out["depth"]), ptr(out["tl"]), ptr(out["br"]), ptr(out["records"]), ptr(out["svec"]),
        ptr(out["alpha"]), ptr(cnt), n_arg, ptr(scratch), scratch.numel(), _stream(mean)),
        "project_cull_fused")
    out["n_dub"] = int(n.value) if sync_count else None
    out["n_dub_dev"] = scratch[:8].view(_I64)
    if not want_projection:
        out["mean2d"] = out["records"][:, 0:2]
        out["cov"] = out["records"][:, 8:12]
    return out


# ---------------------------------------------------------------- a6
def tile_culling_aabb_start_end(aabb_topleft, aabb_bottomright, gaussian_ids, start, end, depth,
                                n_tiles_h, n_tiles_w, sorted_keys=None, check_count=True):
    for t, n in ((aabb_topleft, "aabb_topleft"), (aabb_bottomright, "aabb_bottomright"),
                 (gaussian_ids, "gaussian_ids"), (start, "start"), (end, "end")):
        _chk(t, n, _I32)
    _chk(depth, "depth", _F32)
    if sorted_keys is not None:
        _chk(sorted_keys, "sorted_keys", _I64)
    N = aabb_topleft.size(0)
    n_dub = gaussian_ids.size(0)
    if start.numel() < n_tiles_h * n_tiles_w or end.numel() < n_tiles_h * n_tiles_w:
        raise RuntimeError("start/end must have n_tiles_h * n_tiles_w elements")
    nbytes = capi.lib.gs3d_binning_scratch_bytes(N, n_dub)
    scratch = _scratch(nbytes, depth.device)
    check(capi.lib.gs3d_tile_culling_aabb_start_end(
        N, n_dub, int(n_tiles_h), int(n_tiles_w), ptr(aabb_topleft), ptr(aabb_bottomright),
        ptr(depth), ptr(gaussian_ids), ptr(start), ptr(end), ptr(sorted_keys),
        1 if check_count else 0, ptr(scratch), scratch.numel(), _stream(depth)),
        "tile_culling_aabb_start_end")


def tile_culling_aabb_start_end_capacity(aabb_topleft, aabb_bottomright, gaussian_ids, start, end, depth,
                                         n_tiles_h, n_tiles_w):
    """Binning without a host round trip: gaussian_ids is a buffer of CAPACITY entries.  -> (n_dub int64 [1],
    overflow int32 [1]) device tensors; overflow != 0 means the capacity was too small and the lists are
    truncated (see gs3d_tile_culling_aabb_start_end_capacity)."""
    for t, n in ((aabb_topleft, "aabb_topleft"), (aabb_bottomright, "aabb_bottomright"),
                 (gaussian_ids, "gaussian_ids"), (start, "start"), (end, "end")):
        _chk(t, n, _I32)
    _chk(depth, "depth", _F32)
    N = aabb_topleft.size(0)
    cap = gaussian_ids.size(0)
    if start.numel() < n_tiles_h * n_tiles_w or end.numel() < n_tiles_h * n_tiles_w:
        raise RuntimeError("start/end must have n_tiles_h * n_tiles_w elements")
    dev = depth.device
    n_dub = torch.empty(1, dtype=_I64, device=dev)
    overflow = torch.empty(1, dtype=_I32, device=dev)
    scratch = _scratch(capi.lib.gs3d_binning_scratch_bytes(N, cap), dev)
    check(capi.lib.gs3d_tile_culling_aabb_start_end_capacity(
        N, cap, int(n_tiles_h), int(n_tiles_w), ptr(aabb_topleft), ptr(aabb_bottomright), ptr(depth),
        ptr(gaussian_ids), ptr(start), ptr(end), ptr(n_dub), ptr(overflow), ptr(scratch), scratch.numel(),
        _stream(depth)), "tile_culling_aabb_start_end_capacity")
    return n_dub, overflow


# ---------------------------------------------------------------- records
def pack_records(mean2d, cov, alpha, depth=None):
    _chk(mean2d, "mean", _F32)
    _chk(cov, "cov", _F32)
    _chk(alpha, "alpha", _F32)
    N = mean2d.size(0)
    rec = torch.empty(N, 12, dtype=_F32, device=mean2d.device)
    check(capi.lib.gs3d_pack_records(N, ptr(mean2d), ptr(cov), ptr(alpha), ptr(depth), ptr(rec),
                                     _stream(mean2d)), "pack_records")
    return rec


def _sh_strides(sh, C):
    """(stride per Gaussian, stride per channel) in floats for an [M,3,>=C*C] tensor whose last
    dimension is dense; lets the fused path pass sh_coeffs[..., :C*C] without a copy."""
    if sh.dim() != 3 or sh.size(1) != 3 or sh.size(2) < C * C or sh.stride(2) != 1:
        raise RuntimeError("sh_coeffs must be [N,3,>=C*C] with a dense last dimension")
    return sh.stride(0), sh.stride(1)


# ---------------------------------------------------------------- a7 / a11
def composite_sh_forward(records, sh, start, end, gaussian_ids, out, topleft, c2w, tile_size,
                         n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C_, thresh,
                         bg_rgb=None, final_T=None, n_contrib=None, exact=True):
    _chk(records, "records", _F32)
    for t, n in ((start, "start"), (end, "end"), (gaussian_ids, "gaussian_ids")):
        _chk(t, n, _I32)
    for t, n in ((out, "out"), (topleft, "topleft"), (c2w, "c2w")):
        _chk(t, n, _F32)
    if not sh.is_cuda or sh.dtype != _F32:
        raise RuntimeError("sh_coeffs must be a CUDA floating tensor")
    sg, sc = _sh_strides(sh, C_)
    if out.numel() < H * W * 3:
        raise RuntimeError("out must have H*W*3 elements")
    check(capi.lib.gs3d_composite_sh_forward(
        records.size(0), ptr(records), ptr(sh), sg, sc, ptr(start), ptr(end), ptr(gaussian_ids),
        ptr(out), ptr(topleft), ptr(c2w), int(tile_size), int(n_tiles_h), int(n_tiles_w),
        float(pixel_size_x), float(pixel_size_y), int(H), int(W), int(C_), float(thresh),
        ptr(bg_rgb), ptr(final_T), ptr(n_contrib), 1 if exact else 0, _stream(out)),
        "tile_based_vol_rendering_sh")


# ---------------------------------------------------------------- a8 / a11
def composite_sh_backward(records, sh, start, end, gaussian_ids, out, grad_out, grad_mean, grad_cov,
                          grad_sh, grad_alpha, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
                          pixel_size_x, pixel_size_y, H, W, C_, thresh, exact=True, peer_ptrs=None,
                          multicast_ptr=None, touched=None):
    """peer_ptrs: device addresses of every rank's grad_sh buffer (this rank included) for the fused
    gradient exchange; multicast_ptr: NVSwitch multicast address of the same buffers (optional);
    touched: uint8 [M], receives 1 for every Gaussian whose gradient rows were written (sparse exchange)."""
    if touched is not None:
        _chk(touched, "touched", _U8)
    _chk(records, "records", _F32)
    for t, n in ((start, "start"), (end, "end"), (gaussian_ids, "gaussian_ids")):
        _chk(t, n, _I32)
    for t, n in ((out, "out"), (grad_out, "grad_out"), (grad_mean, "grad_mean"),
                 (grad_cov, "grad_cov"), (grad_alpha, "grad_alpha"), (topleft, "topleft"),
                 (c2w, "c2w")):
        _chk(t, n, _F32)
    sg, sc = _sh_strides(sh, C_)
    gsg, gsc = _sh_strides(grad_sh, C_)
    # the kernels index every per-Gaussian buffer with the ids of the tile lists (< M): a buffer sized
    # for another M (e.g. gradient views kept across a split) would be written out of bounds
    M = records.size(0)
    for t, n, per in ((sh, "sh_coeffs", None), (grad_sh, "grad_sh_coeffs", None), (grad_mean, "grad_mean", 2),
                      (grad_cov, "grad_cov", 4), (grad_alpha, "grad_alpha", 1), (touched, "touched", 1)):
        if t is None:
            continue
        rows = t.size(0) if per is None else t.numel() // per
        if rows != M or (per is not None and t.numel() != per * M):
            raise RuntimeError(f"{n} holds {rows} rows but the staging records hold {M} Gaussians")
    if out.numel() < H * W * 3 or grad_out.numel() < H * W * 3:
        raise RuntimeError("out / grad_out must have H*W*3 elements")
    n_peers = 0 if not peer_ptrs else len(peer_ptrs)
    arr = (C.c_uint64 * max(n_peers, 1))(*([int(x) for x in peer_ptrs] if n_peers else [0]))
    check(capi.lib.gs3d_composite_sh_backward_peers(
        records.size(0), ptr(records), ptr(sh), sg, sc, ptr(start), ptr(end), ptr(gaussian_ids),
        ptr(out), ptr(grad_out), ptr(grad_mean), ptr(grad_cov), ptr(grad_sh), gsg, gsc,
        ptr(grad_alpha), ptr(topleft), ptr(c2w), int(tile_size), int(n_tiles_h), int(n_tiles_w),
        float(pixel_size_x), float(pixel_size_y), int(H), int(W), int(C_), float(thresh),
        1 if exact else 0, C.cast(arr, C.c_void_p), n_peers,
        C.c_void_p(int(multicast_ptr)) if multicast_ptr else None, ptr(touched), _stream(out)),
        "tile_based_vol_rendering_backward_sh")


# ---------------------------------------------------------------- a9 + a10
def project_backward_fused(mask, mean, qvec, svec_param, alpha_param, svec_act, alpha_act, c2w,
                           detach_depth, g_mean2d, g_cov, g_alpha, grad_mean_acc=None, adc_mode=0,
                           out=None, accumulate=False, sparse_filter=False):
    """`out` = (grad_mean, grad_qvec, grad_svec_param, grad_alpha_param) caller buffers (e.g. views of
    one flat all-reduce buffer); with accumulate=True the kernel adds into them.  sparse_filter=True (with
    accumulate): `mask` marks only a few percent of the rows (the compositing backward's `touched` marks) --
    the kernel compacts them first."""
    N = mean.size(0)
    dev = mean.device
    if out is None:
        gm = torch.empty(N, 3, dtype=_F32, device=dev)
        gq = torch.empty(N, 4, dtype=_F32, device=dev)
        gs = torch.empty(N, 3, dtype=_F32, device=dev)
        ga = torch.empty(N, dtype=_F32, device=dev)
        accumulate = False
    else:
        gm, gq, gs, ga = out
        for t, n, per in ((gm, "grad_mean", 3), (gq, "grad_qvec", 4), (gs, "grad_svec", 3), (ga, "grad_alpha", 1)):
            _chk(t, n, _F32)
            if t.numel() != per * N:
                raise RuntimeError(f"{n} has {t.numel()} elements, expected {per} x N = {per * N}")
    for t, n, per in ((mask, "mask", 1), (qvec, "qvec", 4), (svec_param, "svec", 3), (alpha_param, "alpha", 1),
                      (g_mean2d, "grad_mean2d", 2), (g_cov, "grad_cov", 4), (g_alpha, "grad_alpha2d", 1),
                      (grad_mean_acc, "grad_mean_acc", 1)):
        if t is not None and t.numel() != per * N:
            raise RuntimeError(f"{n} has {t.numel()} elements, expected {per} x N = {per * N}")
    check(capi.lib.gs3d_project_backward_fused(
        N, ptr(mask), ptr(mean), ptr(qvec), ptr(svec_param), ptr(alpha_param), int(svec_act),
        int(alpha_act), ptr(c2w), 1 if detach_depth else 0, ptr(g_mean2d), ptr(g_cov), ptr(g_alpha),
        ptr(gm), ptr(gq), ptr(gs), ptr(ga), ptr(grad_mean_acc), int(adc_mode),
        (2 if sparse_filter else 1) if accumulate else 0,
        _stream(mean)), "project_backward_fused")
    return gm, gq, gs, ga


# ---------------------------------------------------------------- sparse gradient exchange
def _rows_args(blocks):
    n = len(blocks)
    for b in blocks:
        _chk(b, "block", _F32)
    ptrs = (C.c_uint64 * n)(*[b.data_ptr() for b in blocks])
    widths = (C.c_uint32 * n)(*[b.numel() // b.size(0) for b in blocks])
    return n, ptrs, widths


def rows_gather(blocks, row_idx, packed):
    """packed[u, :] = concat_s blocks[s][row_idx[u]].flatten()  (blocks: [N, ...] contiguous FP32)."""
    _chk(row_idx, "row_idx", _I32)
    _chk(packed, "packed", _F32)
    n, ptrs, widths = _rows_args(blocks)
    check(capi.lib.gs3d_rows_gather(n, C.cast(ptrs, C.c_void_p), C.cast(widths, C.c_void_p), ptr(row_idx),
                                    row_idx.numel(), ptr(packed), packed.size(1), _stream(packed)), "rows_gather")
    return packed


def rows_scatter(blocks, row_idx, packed):
    """blocks[s][row_idx[u]] = the matching columns of packed[u, :]."""
    _chk(row_idx, "row_idx", _I32)
    _chk(packed, "packed", _F32)
    n, ptrs, widths = _rows_args(blocks)
    check(capi.lib.gs3d_rows_scatter(n, C.cast(ptrs, C.c_void_p), C.cast(widths, C.c_void_p), ptr(row_idx),
                                     row_idx.numel(), ptr(packed), packed.size(1), _stream(packed)), "rows_scatter")


def rows_push_marked(marks, blocks, dst_offsets, peer_result_ptrs, peer_union_ptrs, multicast_ptr=None):
    """Add the marked rows of the private `blocks` into every rank's result buffer (symmetric memory;
    see gs3d_rows_push_marked).  dst_offsets: float offset of each block inside the result buffer."""
    _chk(marks, "marks", _U8)
    n, ptrs, widths = _rows_args(blocks)
    offs = (C.c_uint64 * n)(*[int(o) for o in dst_offsets])
    npeer = len(peer_result_ptrs)
    res = (C.c_uint64 * npeer)(*[int(x) for x in peer_result_ptrs])
    uni = (C.c_uint64 * npeer)(*[int(x) for x in peer_union_ptrs]) if peer_union_ptrs else None
    check(capi.lib.gs3d_rows_push_marked(
        ptr(marks), marks.numel(), n, C.cast(ptrs, C.c_void_p), C.cast(widths, C.c_void_p),
        C.cast(offs, C.c_void_p), C.cast(res, C.c_void_p), C.cast(uni, C.c_void_p) if uni is not None else None,
        npeer, C.c_void_p(int(multicast_ptr)) if multicast_ptr else None, _stream(marks)), "rows_push_marked")


def rows_zero_marked(marks, blocks, clear_marks=True):
    """Zero the rows of every marked Gaussian in `blocks` (may be empty), then clear the marks."""
    _chk(marks, "marks", _U8)
    if blocks:
        n, ptrs, widths = _rows_args(blocks)
        a, b = C.cast(ptrs, C.c_void_p), C.cast(widths, C.c_void_p)
    else:
        n, a, b = 0, None, None
    check(capi.lib.gs3d_rows_zero_marked(ptr(marks), marks.numel(), n, a, b, 1 if clear_marks else 0,
                                         _stream(marks)), "rows_zero_marked")


def marks_broadcast(marks, peer_union_ptrs, multicast_union=None):
    """Set byte g of every rank's union marks for each g with marks[g] != 0 (multicast_union: NVSwitch
    multicast address of the union marks -> one multimem.red.or per non-zero 4-mark word)."""
    _chk(marks, "marks", _U8)
    n = len(peer_union_ptrs)
    uni = (C.c_uint64 * n)(*[int(x) for x in peer_union_ptrs])
    check(capi.lib.gs3d_marks_broadcast(ptr(marks), marks.numel(), C.cast(uni, C.c_void_p), n,
                                        C.c_void_p(int(multicast_union)) if multicast_union else None,
                                        _stream(marks)), "marks_broadcast")


def rows_pull_marked(union_marks, widths, offsets, peer_private_ptrs, peer_result_ptrs, rank,
                     multicast_private=None, multicast_result=None):
    """Sparse NVLS all-reduce of the union-marked rows (see gs3d_rows_pull_marked)."""
    _chk(union_marks, "union_marks", _U8)
    n, npeer = len(widths), len(peer_private_ptrs)
    w = (C.c_uint32 * n)(*[int(x) for x in widths])
    o = (C.c_uint64 * n)(*[int(x) for x in offsets])
    src = (C.c_uint64 * npeer)(*[int(x) for x in peer_private_ptrs])
    dst = (C.c_uint64 * npeer)(*[int(x) for x in peer_result_ptrs])
    check(capi.lib.gs3d_rows_pull_marked(
        ptr(union_marks), union_marks.numel(), n, C.cast(w, C.c_void_p), C.cast(o, C.c_void_p),
        C.cast(src, C.c_void_p), C.cast(dst, C.c_void_p), npeer, int(rank),
        C.c_void_p(int(multicast_private)) if multicast_private else None,
        C.c_void_p(int(multicast_result)) if multicast_result else None, _stream(union_marks)), "rows_pull_marked")


# ---------------------------------------------------------------- tile-row bands (tile-sharded render)
def row_duplicate_counts(aabb_topleft, aabb_bottomright, n_tiles_h):
    """Duplicates per tile row, int64 [n_tiles_h] (device)."""
    _chk(aabb_topleft, "aabb_topleft", _I32)
    _chk(aabb_bottomright, "aabb_bottomright", _I32)
    dev = aabb_topleft.device
    out = torch.empty(int(n_tiles_h), dtype=_I64, device=dev)
    scratch = _scratch(8 * (int(n_tiles_h) + 1), dev)
    check(capi.lib.gs3d_row_duplicate_counts(aabb_topleft.size(0), ptr(aabb_topleft), ptr(aabb_bottomright),
                                             int(n_tiles_h), ptr(out), ptr(scratch), scratch.numel(),
                                             _stream(aabb_topleft)), "row_duplicate_counts")
    return out


def clip_rects_to_rows(aabb_topleft, aabb_bottomright, depth, row_begin, row_end):
    """-> (tl [M',2], br [M',2], depth [M',1], index int32 [M'], n_dub_band): the Gaussians whose rect
    still covers a tile after clipping to tile rows [row_begin, row_end), in ascending index order."""
    _chk(aabb_topleft, "aabb_topleft", _I32)
    _chk(aabb_bottomright, "aabb_bottomright", _I32)
    _chk(depth, "depth", _F32)
    N = aabb_topleft.size(0)
    dev = aabb_topleft.device
    tl = torch.empty(N, 2, dtype=_I32, device=dev)
    br = torch.empty(N, 2, dtype=_I32, device=dev)
    dp = torch.empty(N, 1, dtype=_F32, device=dev)
    idx = torch.empty(N, dtype=_I32, device=dev)
    counts = (C.c_int64 * 2)(0, 0)
    scratch = _scratch(capi.lib.gs3d_clip_scratch_bytes(N), dev)
    check(capi.lib.gs3d_clip_rects_to_rows(N, ptr(aabb_topleft), ptr(aabb_bottomright), ptr(depth), int(row_begin),
                                           int(row_end), ptr(tl), ptr(br), ptr(dp), ptr(idx),
                                           C.cast(counts, _I64P_t), ptr(scratch), scratch.numel(),
                                           _stream(depth)), "clip_rects_to_rows")
    m = int(counts[0])
    return tl[:m], br[:m], dp[:m], idx[:m], int(counts[1])


# ---------------------------------------------------------------- legacy RGB path (8f rank 2)
def composite_rgb_forward(records, color, start, end, gaussian_ids, out, topleft, tile_size, n_tiles_h,
                          n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh, exact=True):
    """tile_based_vol_rendering_start_end (vol_render.h:716-798) on staging records; color [M,3]."""
    _chk(records, "records", _F32)
    for t, n in ((start, "start"), (end, "end"), (gaussian_ids, "gaussian_ids")):
        _chk(t, n, _I32)
    for t, n in ((color, "color"), (out, "out"), (topleft, "topleft")):
        _chk(t, n, _F32)
    if color.numel() != 3 * records.size(0):
        raise RuntimeError("color must be [M,3]")
    if out.numel() < H * W * 3:
        raise RuntimeError("out must have H*W*3 elements")
    check(capi.lib.gs3d_composite_rgb_forward(
        records.size(0), ptr(records), ptr(color), ptr(start), ptr(end), ptr(gaussian_ids), ptr(out),
        ptr(topleft), int(tile_size), int(n_tiles_h), int(n_tiles_w), float(pixel_size_x), float(pixel_size_y),
        int(H), int(W), float(thresh), 1 if exact else 0, _stream(out)), "tile_based_vol_rendering_start_end")


def composite_rgb_backward(records, color, start, end, gaussian_ids, out, grad_out, grad_mean, grad_cov,
                           grad_color, grad_alpha, topleft, tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                           pixel_size_y, H, W, thresh, exact=True):
    """tile_based_vol_rendering_backward_start_end (vol_render.h:800-923); gradients accumulated."""
    _chk(records, "records", _F32)
    for t, n in ((start, "start"), (end, "end"), (gaussian_ids, "gaussian_ids")):
        _chk(t, n, _I32)
    for t, n in ((color, "color"), (out, "out"), (grad_out, "grad_out"), (grad_mean, "grad_mean"),
                 (grad_cov, "grad_cov"), (grad_color, "grad_color"), (grad_alpha, "grad_alpha"),
                 (topleft, "topleft")):
        _chk(t, n, _F32)
    if color.numel() != 3 * records.size(0) or grad_color.numel() != color.numel():
        raise RuntimeError("color / grad_color must be [M,3]")
    check(capi.lib.gs3d_composite_rgb_backward(
        records.size(0), ptr(records), ptr(color), ptr(start), ptr(end), ptr(gaussian_ids), ptr(out),
        ptr(grad_out), ptr(grad_mean), ptr(grad_cov), ptr(grad_color), ptr(grad_alpha), ptr(topleft),
        int(tile_size), int(n_tiles_h), int(n_tiles_w), float(pixel_size_x), float(pixel_size_y), int(H), int(W),
        float(thresh), 1 if exact else 0, _stream(out)), "tile_based_vol_rendering_backward_start_end")


# ---------------------------------------------------------------- loss step (8f rank 4)
def image_loss(out, gt, base="l2", ssim_mult=0.0, window_size=11, want_grad=True):
    """utils/loss.py:5-24 in one call: -> (loss [1] float32 on the device, d loss / d out [H,W,3] or None)."""
    _chk(out, "out", _F32)
    _chk(gt, "gt", _F32)
    if out.dim() != 3 or out.size(2) != 3 or out.shape != gt.shape:
        raise RuntimeError("image_loss: out and gt must both be [H, W, 3]")
    code = {"l1": 1, "l2": 2}.get(base)
    if code is None:
        raise NotImplementedError(base)
    H, W = out.size(0), out.size(1)
    loss = torch.empty(1, dtype=_F32, device=out.device)
    grad = torch.empty_like(out) if want_grad else None
    scratch = _scratch(capi.lib.gs3d_image_loss_scratch_bytes(H, W), out.device)
    check(capi.lib.gs3d_image_loss(ptr(out), ptr(gt), H, W, code, float(ssim_mult), int(window_size), ptr(loss),
                                   ptr(grad), ptr(scratch), scratch.numel(), _stream(out)), "image_loss")
    return loss, grad


# ---------------------------------------------------------------- measurement aid
def set_stage_counters(counters):
    """counters: int64 [4] CUDA tensor (zeroed) or None; see gs3d_set_stage_counters."""
    if counters is not None:
        _chk(counters, "counters", _I64)
        if counters.numel() < 4:
            raise RuntimeError("counters must hold four int64 values")
    check(capi.lib.gs3d_set_stage_counters(ptr(counters)), "set_stage_counters")


# ---------------------------------------------------------------- optimiser step (8f rank 1)
def adam_step(params, grads, exp_avgs, exp_avg_sqs, lrs, beta1, beta2, eps, step, state_mode=0):
    """One Adam step over up to 8 tensors in one launch (see gs3d_adam_step).  exp_avgs / exp_avg_sqs
    may be None with state_mode 2."""
    n = len(params)
    if not (len(grads) == n and len(lrs) == n):
        raise RuntimeError("adam_step: params / grads / lrs differ in length")
    if n == 0:
        return
    segs = (capi.AdamSegment * n)()
    for i in range(n):
        p, g = _chk(params[i], "param", _F32), _chk(grads[i], "grad", _F32)
        if g.numel() != p.numel() or g.device != p.device:
            raise RuntimeError("adam_step: grad does not match its parameter")
        m = v = None
        if state_mode != 2:
            m, v = _chk(exp_avgs[i], "exp_avg", _F32), _chk(exp_avg_sqs[i], "exp_avg_sq", _F32)
            if m.numel() != p.numel() or v.numel() != p.numel():
                raise RuntimeError("adam_step: moment buffer does not match its parameter")
        segs[i] = capi.AdamSegment(p.data_ptr(), g.data_ptr(), m.data_ptr() if m is not None else None,
                                   v.data_ptr() if v is not None else None, p.numel(), float(lrs[i]))
    check(capi.lib.gs3d_adam_step(n, segs, float(beta1), float(beta2), float(eps), int(step), int(state_mode),
                                  _stream(params[0])), "adam_step")


# ---------------------------------------------------------------- adaptive density control (8f rank 1)
ADC_KEEP, ADC_CLONE, ADC_SPLIT, ADC_DROP = 0, 1, 2, 3


def adc_classify(grad_mean_acc, cnt, reduction, pos_grad_thresh, svec_param, svec_act, split_scale_thresh):
    """-> cls uint8 [N] (KEEP / CLONE / SPLIT), sh_renderer.py:433-456.  reduction: "max" | "mean"."""
    _chk(grad_mean_acc, "grad_mean", _F32)
    _chk(svec_param, "svec_before_activation", _F32)
    red = {"max": 1, "mean": 2}.get(reduction)
    if red is None:
        raise NotImplementedError(reduction)
    if red == 2:
        _chk(cnt, "cnt", _I32)
    N = grad_mean_acc.numel()
    cls = torch.empty(N, dtype=_U8, device=grad_mean_acc.device)
    check(capi.lib.gs3d_adc_classify(N, ptr(grad_mean_acc), ptr(cnt) if red == 2 else None, red,
                                     float(pos_grad_thresh), ptr(svec_param), int(svec_act),
                                     float(split_scale_thresh), ptr(cls), _stream(cls)), "adc_classify")
    return cls


def adc_classify_alpha(alpha_param, alpha_act, alpha_thresh):
    """-> cls uint8 [N]: KEEP iff act(alpha) >= thresh else DROP (sh_renderer.py:542-543)."""
    _chk(alpha_param, "alpha_before_activation", _F32)
    N = alpha_param.numel()
    cls = torch.empty(N, dtype=_U8, device=alpha_param.device)
    check(capi.lib.gs3d_adc_classify_alpha(N, ptr(alpha_param), int(alpha_act), float(alpha_thresh), ptr(cls),
                                           _stream(cls)), "adc_classify_alpha")
    return cls


def adc_plan(cls):
    """cls: uint8 classes or a torch.bool keep mask.  -> (plan, (n_stay, n_clone, n_split)); syncs."""
    _chk(cls, "cls", (_U8, _BOOL))
    N = cls.numel()
    scratch = _scratch(capi.lib.gs3d_adc_scratch_bytes(N), cls.device)
    counts = (C.c_int64 * 3)(0, 0, 0)
    check(capi.lib.gs3d_adc_plan(N, ptr(cls), 1 if cls.dtype == _BOOL else 0, C.cast(counts, _I64P_t),
                                 ptr(scratch), scratch.numel(), _stream(cls)), "adc_plan")
    return (cls, scratch, counts), (int(counts[0]), int(counts[1]), int(counts[2]))


def adc_apply(plan, mean, qvec, svec_param, sh_coeffs, alpha_param, svec_act=1, scale_shrink_factor=1.0,
              noise=None):
    """-> the five new parameter tensors (see gs3d_adc_apply)."""
    cls, scratch, counts = plan
    for t, n in ((mean, "mean"), (qvec, "qvec"), (svec_param, "svec_before_activation"),
                 (sh_coeffs, "sh_coeffs"), (alpha_param, "alpha_before_activation")):
        _chk(t, n, _F32)
    N = mean.size(0)
    if cls.numel() != N:
        raise RuntimeError("adc_apply: plan was made for a different N")
    n_out = int(counts[0]) + int(counts[1]) + 2 * int(counts[2])
    if int(counts[2]):
        _chk(noise, "noise", _F32)
        if noise.numel() != 6 * int(counts[2]):
            raise RuntimeError("adc_apply: noise must be [2*n_split, 3]")
    dev = mean.device
    sh_width = sh_coeffs[0].numel() if N else 0
    out = (torch.empty(n_out, 3, dtype=_F32, device=dev), torch.empty(n_out, 4, dtype=_F32, device=dev),
           torch.empty(n_out, 3, dtype=_F32, device=dev),
           torch.empty((n_out,) + tuple(sh_coeffs.shape[1:]), dtype=_F32, device=dev),
           torch.empty(n_out, dtype=_F32, device=dev))
    check(capi.lib.gs3d_adc_apply(N, ptr(cls), 1 if cls.dtype == _BOOL else 0, ptr(scratch),
                                  C.cast(counts, _I64P_t), ptr(mean), ptr(qvec), ptr(svec_param), ptr(sh_coeffs),
                                  ptr(alpha_param), int(sh_width), int(svec_act), float(scale_shrink_factor),
                                  ptr(noise) if noise is not None else None, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                  ptr(out[3]), ptr(out[4]), _stream(mean)), "adc_apply")
    return out


# ---------------------------------------------------------------- device guard
# The kernels are launched on the CUDA *current* device with the stream of the tensors' device.  A model
# that lives on another device than the current one (cfg.device = "cuda:1" without torch.cuda.set_device)
# would launch on the wrong GPU: run such calls under that device's context (the reference has no guard
# either -- render.cu:12 includes CUDAGuard.h and never uses it -- but there the mistake is silent).
def _first_cuda_tensor(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a
        elif isinstance(a, (list, tuple)):
            for b in a:
                if isinstance(b, torch.Tensor) and b.is_cuda:
                    return b
    return None


def _device_guarded(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        t = _first_cuda_tensor(args, kwargs)
        if t is not None and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return fn(*args, **kwargs)
        return fn(*args, **kwargs)

    return wrapped


for _name, _fn in list(globals().items()):
    if callable(_fn) and getattr(_fn, "__module__", None) == __name__ and not _name.startswith("_") \
            and _name not in ("check", "ptr"):
        globals()[_name] = _device_guarded(_fn)
del _name, _fn
