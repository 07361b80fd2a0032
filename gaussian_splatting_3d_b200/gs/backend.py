"""`gs/backend.py` of the reference JIT-compiles its CUDA extension and exposes it as `_backend`
(backend.py:52-67).  Here `_backend` is the `_gs`-compatible module over libgs3d_b200.so, built
ahead of time by `__graft_entry__.build()` / `_build.build()`; importing fails loudly if the
library is missing (no CPU fallback)."""
from .. import _gs as _backend

__all__ = ["_backend"]
