"""ctypes binding of libgs3d_b200.so -- the C ABI declared in include/gs3d_b200.h.

This is the ONLY route from Python to the kernels.  There is no CPU fallback: if the shared
library is missing or cannot be loaded the import raises, and every wrapper raises RuntimeError
(with gs3d_last_error()) on a non-zero return code, mirroring TORCH_CHECK -> RuntimeError in the
reference's bindings (gs/src/include/common.h:29-54).
"""
import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# GS3D_LIB: developer override used by tools/ablate.sh to time experimental builds of the same ABI
LIB_PATH = Path(os.environ["GS3D_LIB"]) if os.environ.get("GS3D_LIB") else _PKG / "libgs3d_b200.so"


class Camera(C.Structure):
    """struct gs3d_camera (mirrors utils/camera.py:219-230 CameraInfo)."""

    _fields_ = [
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("w", C.c_int32), ("h", C.c_int32),
        ("near_plane", C.c_double), ("far_plane", C.c_double),
    ]


class AdamSegment(C.Structure):
    """struct gs3d_adam_segment: one parameter tensor of the optimiser step."""

    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("n", C.c_uint64), ("lr", C.c_double),
    ]


_P = C.c_void_p
_u32 = C.c_uint32
_f = C.c_float
_d = C.c_double
_i = C.c_int
_sz = C.c_size_t
_CAM = C.POINTER(Camera)
_I64P = C.POINTER(C.c_int64)
_ADAM = C.POINTER(AdamSegment)

# name -> (restype, argtypes); must list every symbol of include/gs3d_b200.h (tests check this)
SIGNATURES = {
    "gs3d_version": (_i, []),
    "gs3d_last_error": (C.c_char_p, []),
    "gs3d_launch_count": (C.c_uint64, []),
    "gs3d_get_frustum": (_i, [_P, _CAM, _P, _P, _P]),
    "gs3d_culling_gaussian_bsphere": (_i, [_u32, _P, _P, _P, _P, _P, _P, _f, _P]),
    "gs3d_project_gaussians": (_i, [_u32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gs3d_project_gaussians_backward": (_i, [_u32, _P, _P, _P, _P, _P, _P, _P, _i, _P, _P, _P, _P]),
    "gs3d_count_scratch_bytes": (_sz, [_u32]),
    "gs3d_tile_culling_aabb_count": (_i, [_u32, _P, _P, _u32, _CAM, _f, _P, _P, _I64P, _P, _sz, _P]),
    "gs3d_project_cull_fused": (_i, [_u32, _P, _P, _P, _P, _i, _i, _P, _CAM, _f, _i, _f, _u32,
                                     _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _sz, _P]),
    "gs3d_binning_scratch_bytes": (_sz, [_u32, _u32]),
    "gs3d_tile_culling_aabb_start_end": (_i, [_u32, _u32, _u32, _u32, _P, _P, _P, _P, _P, _P, _P,
                                              _i, _P, _sz, _P]),
    "gs3d_tile_culling_aabb_start_end_capacity": (_i, [_u32, _u32, _u32, _u32, _P, _P, _P, _P, _P, _P, _P, _P,
                                                       _P, _sz, _P]),
    "gs3d_pack_records": (_i, [_u32, _P, _P, _P, _P, _P, _P]),
    "gs3d_composite_sh_forward": (_i, [_u32, _P, _P, _u32, _u32, _P, _P, _P, _P, _P, _P, _u32, _u32,
                                       _u32, _f, _f, _u32, _u32, _u32, _f, _P, _P, _P, _i, _P]),
    "gs3d_composite_sh_backward": (_i, [_u32, _P, _P, _u32, _u32, _P, _P, _P, _P, _P, _P, _P, _P,
                                        _u32, _u32, _P, _P, _P, _u32, _u32, _u32, _f, _f, _u32, _u32,
                                        _u32, _f, _i, _P]),
    "gs3d_composite_sh_backward_peers": (_i, [_u32, _P, _P, _u32, _u32, _P, _P, _P, _P, _P, _P, _P, _P,
                                              _u32, _u32, _P, _P, _P, _u32, _u32, _u32, _f, _f, _u32, _u32,
                                              _u32, _f, _i, _P, _i, _P, _P, _P]),
    "gs3d_rows_gather": (_i, [_i, _P, _P, _P, _u32, _P, _u32, _P]),
    "gs3d_rows_scatter": (_i, [_i, _P, _P, _P, _u32, _P, _u32, _P]),
    "gs3d_rows_push_marked": (_i, [_P, _u32, _i, _P, _P, _P, _P, _P, _i, _P, _P]),
    "gs3d_rows_zero_marked": (_i, [_P, _u32, _i, _P, _P, _i, _P]),
    "gs3d_marks_broadcast": (_i, [_P, _u32, _P, _i, _P, _P]),
    "gs3d_row_duplicate_counts": (_i, [_u32, _P, _P, _u32, _P, _P, _sz, _P]),
    "gs3d_clip_scratch_bytes": (_sz, [_u32]),
    "gs3d_clip_rects_to_rows": (_i, [_u32, _P, _P, _P, _i, _i, _P, _P, _P, _P, _I64P, _P, _sz, _P]),
    "gs3d_rows_pull_marked": (_i, [_P, _u32, _i, _P, _P, _P, _P, _i, _i, _P, _P, _P]),
    "gs3d_project_backward_fused": (_i, [_u32, _P, _P, _P, _P, _P, _i, _i, _P, _i, _P, _P, _P, _P, _P,
                                         _P, _P, _P, _i, _i, _P]),
    "gs3d_composite_rgb_forward": (_i, [_u32, _P, _P, _P, _P, _P, _P, _P, _u32, _u32, _u32, _f, _f, _u32, _u32,
                                        _f, _i, _P]),
    "gs3d_composite_rgb_backward": (_i, [_u32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _u32, _u32, _u32,
                                         _f, _f, _u32, _u32, _f, _i, _P]),
    "gs3d_image_loss_scratch_bytes": (_sz, [_u32, _u32]),
    "gs3d_image_loss": (_i, [_P, _P, _u32, _u32, _i, _f, _u32, _P, _P, _P, _sz, _P]),
    "gs3d_set_stage_counters": (_i, [_P]),
    "gs3d_adam_step": (_i, [_i, _ADAM, _d, _d, _d, _u32, _i, _P]),
    "gs3d_adc_classify": (_i, [_u32, _P, _P, _i, _f, _P, _i, _f, _P, _P]),
    "gs3d_adc_classify_alpha": (_i, [_u32, _P, _i, _f, _P, _P]),
    "gs3d_adc_scratch_bytes": (_sz, [_u32]),
    "gs3d_adc_plan": (_i, [_u32, _P, _i, _I64P, _P, _sz, _P]),
    "gs3d_adc_apply": (_i, [_u32, _P, _i, _P, _I64P, _P, _P, _P, _P, _P, _u32, _i, _f, _P, _P, _P, _P, _P, _P, _P]),
}


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class GS3DError(RuntimeError):
    pass


def last_error():
    return lib.gs3d_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise GS3DError(f"{what or 'gs3d call'} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device/host address of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream(device=None):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def camera_struct(camera_info):
    return Camera(
        float(camera_info.fx), float(camera_info.fy), float(camera_info.cx), float(camera_info.cy),
        int(camera_info.w), int(camera_info.h),
        float(camera_info.near_plane), float(camera_info.far_plane),
    )
