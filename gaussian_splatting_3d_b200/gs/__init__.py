"""Mirror of the reference's `gs` package API (gs/renderer.py, gs/sh_renderer.py, gs/culling.py,
gs/backend.py) on top of the B200-native kernels."""
