"""B200-native Gaussian-splatting rasteriser: drop-in for the hot path of
heheyas/gaussian_splatting_3d (cull -> project -> tile-bin/sort -> alpha-composite with SH colour
-> backward) behind the reference's own `_gs` / `gs.*` API.

    import gaussian_splatting_3d_b200 as g3
    g3.install()            # registers `_gs`, `gs`, `gs.renderer`, `gs.sh_renderer`, `gs.culling`,
                            # `gs.backend`, `utils.camera`, ... in sys.modules
    from gs.sh_renderer import SHRenderer      # the reference's import lines now resolve here

The compute lives in libgs3d_b200.so (hand-written sm_100a CUDA, C ABI in include/gs3d_b200.h).
There is no CPU fallback: importing `gaussian_splatting_3d_b200.capi` raises if the library has
not been built.
"""
import importlib
import sys

__version__ = "0.1.0"

_ALIASES = {
    "_gs": "gaussian_splatting_3d_b200._gs",
    "gs": "gaussian_splatting_3d_b200.gs",
    "gs.backend": "gaussian_splatting_3d_b200.gs.backend",
    "gs.culling": "gaussian_splatting_3d_b200.gs.culling",
    "gs.renderer": "gaussian_splatting_3d_b200.gs.renderer",
    "gs.sh_renderer": "gaussian_splatting_3d_b200.gs.sh_renderer",
    "utils": "gaussian_splatting_3d_b200.utils",
    "utils.camera": "gaussian_splatting_3d_b200.utils.camera",
    "utils.transforms": "gaussian_splatting_3d_b200.utils.transforms",
    "utils.activations": "gaussian_splatting_3d_b200.utils.activations",
    "utils.schedulers": "gaussian_splatting_3d_b200.utils.schedulers",
    "utils.misc": "gaussian_splatting_3d_b200.utils.misc",
    "utils.loss": "gaussian_splatting_3d_b200.utils.loss",
}


def install(overwrite=False):
    """Make the reference's module names resolve to this package."""
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not overwrite:
            continue
        sys.modules[alias] = importlib.import_module(target)
    return sys.modules["gs"]
